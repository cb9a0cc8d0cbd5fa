"""BASELINE.json config 4 on one GPU: an input larger than 4 GiB (64-bit offsets everywhere), level 6, device-resident
deflate + inflate, bit-exact, Adler-32 against libz pieces folded with the combine identity.
usage: big_check.py [gib=5]"""
import os, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import zlibk_b200
from zlibk_b200 import corpus

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(gib * (1 << 30)) + 12345
eng = zlibk_b200.Engine(0)
dev = torch.device("cuda", 0)
piece = 1 << 30
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
want = 1
for off in range(0, n, piece):
    ln = min(piece, n - off)
    host = corpus.mixed_entropy(ln, seed_offset=off >> 30) if "seed_offset" in corpus.mixed_entropy.__code__.co_varnames else np.roll(corpus.mixed_entropy(ln), off >> 20)
    d_in[off:off + ln] = torch.from_numpy(host).to(dev)
    want = eng.adler32_combine(want, zlib.adler32(host), ln)
cap = eng.compress_bound(n)
d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
d_res = torch.zeros(4, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
assert eng.deflate_device(d_in.data_ptr(), n, 6, 0, d_out.data_ptr(), cap, d_res.data_ptr()) == 0
eng.synchronize()
t1 = time.perf_counter()
out_len, adler, status, _ = d_res.cpu().tolist()
assert status == 0 and adler == want, (status, adler, want)
d_back = torch.empty(n, dtype=torch.uint8, device=dev)
d_res2 = torch.zeros(4, dtype=torch.int64, device=dev)
t2 = time.perf_counter()
assert eng.inflate_device(d_out.data_ptr(), out_len, d_back.data_ptr(), n, d_res2.data_ptr()) == 0
eng.synchronize()
t3 = time.perf_counter()
r = d_res2.cpu().tolist()
assert r[2] == 1 and r[0] == n and r[1] == out_len, r
assert bool(torch.equal(d_back, d_in))
print("big ok: %d bytes -> %d (%.4f); deflate %.1f GB/s, inflate %.1f GB/s (wall clock, first call incl. scratch allocation); peak device memory %.1f GiB" % (
    n, out_len, out_len / n, n / (t1 - t0) / 1e9, n / (t3 - t2) / 1e9, torch.cuda.max_memory_allocated() / 2 ** 30))
