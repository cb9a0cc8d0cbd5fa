"""BASELINE.json config 3: level sweep (1..9) of deflate throughput and compressed size on the synthetic corpora, with the
reference restatement (oracle) and libz sizes on a bounded prefix beside them.
usage: level_sweep.py [size_mib=1024] [prefix_mib=16]"""
import os, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import zlibk_b200
from zlibk_b200 import corpus
from tests.oracle_binding import Oracle

size_mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prefix_mib = int(sys.argv[2]) if len(sys.argv) > 2 else 16
n, pn = size_mib << 20, prefix_mib << 20
oracle = Oracle(os.path.join(ROOT, "oracle", "libzk_oracle.so"))
eng = zlibk_b200.Engine(0)
dev = torch.device("cuda", 0)
print("| corpus | level | GPU deflate GB/s | GPU inflate GB/s | GPU size/input | GPU size on prefix | reference (oracle) on prefix | libz on prefix |")
print("|---|---:|---:|---:|---:|---:|---:|---:|")
for kind in ("text", "mixed"):
    host = corpus.text_like(n) if kind == "text" else corpus.mixed_entropy(n)
    d_in = torch.from_numpy(host).to(dev)
    cap = eng.compress_bound(n)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_back = torch.empty(n, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(4, dtype=torch.int64, device=dev)
    d_res2 = torch.zeros(4, dtype=torch.int64, device=dev)
    prefix = host[:pn].tobytes()
    for level in (1, 2, 3, 4, 5, 6, 7, 8, 9):
        best = 1e9
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc = eng.deflate_device(d_in.data_ptr(), n, level, 0, d_out.data_ptr(), cap, d_res.data_ptr())
            eng.synchronize()
            best = min(best, time.perf_counter() - t0)
        m = d_res.cpu().tolist()[0]
        ibest = 1e9
        for it in range(2):
            t0 = time.perf_counter()
            eng.inflate_device(d_out.data_ptr(), m, d_back.data_ptr(), n, d_res2.data_ptr())
            eng.synchronize()
            ibest = min(ibest, time.perf_counter() - t0)
        assert d_res2.cpu().tolist()[2] == 1 and bool(torch.equal(d_back, d_in))
        rc, s, _ = eng.deflate(prefix, level)
        gp = len(s) / pn
        ref = lz = ""
        if level in (1, 6, 9):
            _, rs, _ = oracle.deflate(prefix, level)
            ref = "%.4f" % (len(rs) / pn)
        lz = "%.4f" % (len(zlib.compress(prefix, level)) / pn)
        print("| %s %d MiB | %d | %.1f | %.1f | %.4f | %.4f | %s | %s |" % (kind, size_mib, level, n / best / 1e9, n / ibest / 1e9, m / n, gp, ref, lz), flush=True)
