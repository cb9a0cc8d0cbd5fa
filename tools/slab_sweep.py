"""Device-resident deflate (zk_deflate_zlib_device) with the input cut into slabs that overlap on two streams.
usage: python tools/slab_sweep.py [mib=1024] [level=6] [reps=7]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import zlibk_b200 as zk
from zlibk_b200 import corpus

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 7
n = mib << 20
dev = torch.device("cuda", 0)
eng = zk.Engine(0)
d_in = torch.from_numpy(corpus.mixed_entropy(n)).to(dev)
cap = eng.compress_bound(n)
d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
d_res = torch.zeros(4, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
M = 1 << 20


def run(label, **opts):
    for k, v in opts.items():
        eng.set_option(k, v)
    ts = []
    for i in range(reps + 2):
        flush.fill_(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        assert eng.deflate_device(d_in.data_ptr(), n, level, 0, d_out.data_ptr(), cap, d_res.data_ptr()) == 0
        eng.synchronize()
        ts.append(time.perf_counter() - t0)
    r = d_res.cpu().tolist()
    assert r[2] == 0, r
    ts = sorted(ts[2:])
    print("%-36s best %6.2f  median %6.2f GB/s   out %d" % (label, n / ts[0] / 1e9, n / ts[len(ts) // 2] / 1e9, r[0]), flush=True)
    for k in opts:
        eng.set_option(k, 1 if k == "slab_overlap" else 0)


run("default (two equal slabs from 512 MiB)")
run("one slab", slab_bytes=1024 * M)
for slab in (896, 832, 768, 704, 640, 576, 512, 384, 256, 192, 128, 96, 64):
    if slab * M >= n:
        continue
    run("slab %d MiB" % slab, slab_bytes=slab * M)
run("slab 256 MiB, overlap off", slab_bytes=256 * M, slab_overlap=0)
for unit in (256, 512):
    run("slab 128 MiB, unit %d KiB" % unit, slab_bytes=128 * M, chunk_bytes=unit << 10)
