"""First-light checks on a GPU box: prints details instead of asserting, so one gpurun call tells as
much as possible.  Not a test and not a benchmark."""
import os
import sys
import time
import traceback
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zlibk_b200
from zlibk_b200 import corpus
from tests.oracle_binding import Oracle

oracle = Oracle(os.path.join(ROOT, "oracle", "libzk_oracle.so"))
eng = zlibk_b200.Engine(0)
print("version", zlibk_b200.version(), "devices", zlibk_b200.library().zk_device_count())


def step(name, fn):
    t = time.time()
    try:
        r = fn()
        print("[%s] %s (%.3fs)" % ("ok" if r in (None, True) else "FAIL", name, time.time() - t), "" if r in (None, True) else r)
    except Exception:
        print("[EXC] %s" % name)
        traceback.print_exc()
    sys.stdout.flush()


def adler_case(n, seed=1):
    data = corpus._random_bytes(n, seed).tobytes()
    got, want = eng.adler32(data), zlib.adler32(data)
    return True if got == want else "n=%d got %08x want %08x" % (n, got, want)


for n in (1, 15, 16, 17, 4095, 5552, 65536, 1000003, 16 << 20):
    step("adler n=%d" % n, lambda n=n: adler_case(n))


def deflate_case(name, data, level):
    rc, stream, adler = eng.deflate(data, level)
    if rc != 0:
        return "rc=%d err=%s" % (rc, eng.last_error())
    stream = bytes(stream)
    try:
        back = zlib.decompress(stream)
    except Exception as e:
        return "libz: %r (len %d -> %d) head=%s" % (e, len(data), len(stream), stream[:24].hex())
    if back != data:
        bad = next(i for i in range(min(len(back), len(data))) if back[i] != data[i]) if len(back) == len(data) else -1
        return "libz output differs (len %d vs %d, first diff %d)" % (len(back), len(data), bad)
    orc, out, _, _ = oracle.inflate(stream, cap=len(data) + 64)
    if orc != 1 or out != data:
        return "oracle inflate rc=%d" % orc
    if adler != zlib.adler32(data):
        return "adler %08x" % adler
    print("    %s L%d: %d -> %d (%.4f; libz6 %.4f)" % (name, level, len(data), len(stream), len(stream) / max(1, len(data)), len(zlib.compress(data, 6)) / max(1, len(data))))
    return True


for name, data in corpus.edge_cases():
    for level in (0, 1, 6, 9):
        step("deflate %s L%d" % (name, level), lambda d=data, l=level, nm=name: deflate_case(nm, d, l))

big_text = corpus.text_like(8 << 20).tobytes()
big_mixed = corpus.mixed_entropy(12 << 20).tobytes()
for level in (1, 6, 9):
    step("deflate text8M L%d" % level, lambda l=level: deflate_case("text8M", big_text, l))
    step("deflate mixed12M L%d" % level, lambda l=level: deflate_case("mixed12M", big_mixed, l))


def inflate_case(name, stream, want_rc, want_out):
    rc, out, used = eng.inflate(stream, max(len(want_out), 1) + 64)
    out = bytes(out)
    if rc != want_rc or out != want_out:
        return "%s: rc=%d want %d, len %d want %d, used %d of %d" % (name, rc, want_rc, len(out), len(want_out), used, len(stream))
    return True


for lvl in (0, 1, 6, 9):
    for nm, d in (("text", big_text[:3000000]), ("mixed", big_mixed[:5000000]), ("small", b"hello world")):
        s = zlib.compress(d, lvl)
        step("inflate libz %s L%d" % (nm, lvl), lambda s=s, d=d, nm=nm: inflate_case(nm, s, 1, d))
for nm, d in (("hello", b"hello world"), ("stale", b"abcabc\0\0\0"), ("text300k", big_text[:300000]), ("aaa", b"a" * 1000)):
    for lvl in (1, 6):
        _, s, _ = oracle.deflate(d, lvl)
        orc, oout, _, _ = oracle.inflate(s, cap=len(d) * 2 + 1000)
        step("inflate reference-stream %s L%d (oracle rc %d)" % (nm, lvl, orc), lambda s=s, orc=orc, oout=oout, nm=nm: inflate_case(nm, s, orc, oout))


def own_roundtrip(data, level):
    rc, stream, _ = eng.deflate(data, level)
    stream = bytes(stream)
    t = time.time()
    rc2, out, used = eng.inflate(stream, len(data))
    dt = time.time() - t
    ok = rc2 == 1 and bytes(out) == data and used == len(stream)
    print("    own round trip L%d: rc=%d ok=%s inflate %.3fs" % (level, rc2, ok, dt))
    return True if ok else "rc=%d len=%d used=%d/%d" % (rc2, len(out), used, len(stream))


for level in (1, 6):
    step("own round trip text8M L%d" % level, lambda l=level: own_roundtrip(big_text, l))
    step("own round trip mixed12M L%d" % level, lambda l=level: own_roundtrip(big_mixed, l))

# rough timing (host buffers, includes copies)
import numpy as np

for nm, d in (("text64M", corpus.text_like(64 << 20)), ("mixed64M", corpus.mixed_entropy(64 << 20))):
    out = np.empty(eng.compress_bound(d.size), dtype=np.uint8)
    for it in range(3):
        t = time.time()
        rc, s, _ = eng.deflate(d, 6, out=out)
        dt = time.time() - t
        print("    deflate %s L6 iter %d: rc=%d %.1f ms -> %.2f GB/s e2e, ratio %.4f" % (nm, it, rc, dt * 1e3, d.size / dt / 1e9, len(s) / d.size))
print("launches", eng.counter("kernel_launches"))
