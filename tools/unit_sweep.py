"""Device-resident deflate of 1 GiB with different LZ unit sizes (option chunk_bytes), default slabs.
usage: python tools/unit_sweep.py [mib=1024] [level=6]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import zlibk_b200 as zk
from zlibk_b200 import corpus

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
n = mib << 20
dev = torch.device("cuda", 0)
eng = zk.Engine(0)
d_in = torch.from_numpy(corpus.mixed_entropy(n)).to(dev)
cap = eng.compress_bound(n)
d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
d_res = torch.zeros(4, dtype=torch.int64, device=dev)
for unit in ([int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else (0, 256, 320, 384, 448, 512, 640, 768, 1024)):
    eng.set_option("chunk_bytes", unit << 10)
    ts = []
    for i in range(7):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        assert eng.deflate_device(d_in.data_ptr(), n, level, 0, d_out.data_ptr(), cap, d_res.data_ptr()) == 0
        eng.synchronize()
        ts.append(time.perf_counter() - t0)
    r = d_res.cpu().tolist()
    ts = sorted(ts[2:])
    print("unit %4s KiB: best %6.2f median %6.2f GB/s, out %d" % (unit or "auto", n / ts[0] / 1e9, n / ts[len(ts) // 2] / 1e9, r[0]), flush=True)
eng.set_option("chunk_bytes", 0)
