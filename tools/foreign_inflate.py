"""Inflate of streams this engine did not make, device-resident: libz one-shot streams (no entry points at all), pigz-style
streams (a sync flush every 128 KiB of input) and the reference's own streams (oracle restatement, small).
usage: python tools/foreign_inflate.py [mib=256] [reps=5] [levels=1,6,9]"""
import json
import os
import sys
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import zlibk_b200 as zk
from zlibk_b200 import corpus

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
levels = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [6]
n = mib << 20
dev = torch.device("cuda", 0)
eng = zk.Engine(0)
eng.set_option("timing", 1)


def libz_stream(data, level, piece=32 << 20):
    """ONE zlib stream equal in content to zlib.compress(data, level) block-wise: pieces are deflated in parallel as raw
    bodies primed with the 32 KiB before them and joined -- here only to make a big stream quickly; every block in it was
    made by libz and there is no marker except at the piece joins (one per 32 MiB)."""
    view = memoryview(data)

    def one(i):
        lo = i * piece
        hi = min(len(data), lo + piece)
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, zlib.Z_DEFAULT_STRATEGY, bytes(view[max(0, lo - 32768): lo])) if lo else \
            zlib.compressobj(level, zlib.DEFLATED, -15)
        body = co.compress(view[lo:hi])
        body += co.flush(zlib.Z_FINISH if hi == len(data) else zlib.Z_SYNC_FLUSH)
        return body

    with ThreadPoolExecutor(os.cpu_count() or 8) as ex:
        bodies = list(ex.map(one, range((len(data) + piece - 1) // piece)))
    return b"\x78\x9c" + b"".join(bodies) + zlib.adler32(data).to_bytes(4, "big")


def pigz_stream(data, level, piece=128 << 10):
    co = zlib.compressobj(level)
    out = []
    for lo in range(0, len(data), piece):
        out.append(co.compress(data[lo: lo + piece]))
        out.append(co.flush(zlib.Z_SYNC_FLUSH))
    out.append(co.flush())
    return b"".join(out)


def run(label, stream, want):
    m = len(stream)
    d_in = torch.from_numpy(np.frombuffer(stream, dtype=np.uint8).copy()).to(dev)
    d_out = torch.empty(len(want) + 64, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(4, dtype=torch.int64, device=dev)
    ts = []
    for i in range(reps + 1):
        torch.cuda.synchronize()
        eng.counter("reset_timing")
        t0 = time.perf_counter()
        rc = eng.inflate_device(d_in.data_ptr(), m, d_out.data_ptr(), len(want), d_res.data_ptr())
        eng.synchronize()
        ts.append(time.perf_counter() - t0)
        assert rc == 0, rc
    res = d_res.cpu().tolist()
    ok = res[2] == 1 and res[0] == len(want) and bytes(d_out[: len(want)].cpu().numpy()) == want
    best = min(ts[1:])
    rec = {"stream": label, "compressed_bytes": m, "bytes": len(want), "ms": round(best * 1e3, 2), "gbps": round(len(want) / best / 1e9, 2),
           "decode_ms": round(eng.counter("inflate_decode_us") / 1e3, 2), "bit_exact": bool(ok), "status": res[2]}
    print(json.dumps(rec), flush=True)
    return rec


for kind in ("mixed", "text"):
    host = corpus.mixed_entropy(n) if kind == "mixed" else corpus.text_like(n)
    want = host.tobytes()
    for level in levels:
        t0 = time.perf_counter()
        s = libz_stream(want, level)
        print("# libz level %d %s: %d -> %d bytes in %.1f s" % (level, kind, n, len(s), time.perf_counter() - t0), flush=True)
        assert zlib.decompress(s) == want
        run("libz-%d %s" % (level, kind), s, want)
    p = pigz_stream(want[: min(n, 128 << 20)], 6)
    run("pigz-style-6 %s (128 KiB pieces)" % kind, p, want[: min(n, 128 << 20)])
