"""Sweep of the pipelined host call (zk_deflate_zlib with pinned host buffers): slab sizes, slab overlap on / off, unit
size.  Prints one line per setting: GB/s of input (best and median of `reps` wall-clock timings), stream size.
usage: python tools/e2e_sweep.py [mib=1024] [level=6] [reps=5]"""
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import zlibk_b200 as zk
from zlibk_b200 import corpus

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
n = mib << 20
eng = zk.Engine(0)
h_in = torch.from_numpy(corpus.mixed_entropy(n)).pin_memory()
h_out = torch.empty(eng.compress_bound(n), dtype=torch.uint8).pin_memory()
adler = zlib.adler32(h_in.numpy())


def run(label, **opts):
    for k, v in opts.items():
        eng.set_option(k, v)
    ts = []
    for i in range(reps + 1):
        t0 = time.perf_counter()
        rc, s, a = eng.deflate(h_in, level, out=h_out)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
        assert rc == 0 and a == adler, (rc, eng.last_error())
    ts = sorted(ts[1:])
    print("%-44s best %6.2f  median %6.2f GB/s   out %d" % (label, n / ts[0] / 1e9, n / ts[len(ts) // 2] / 1e9, len(s)), flush=True)
    for k in opts:
        eng.set_option(k, 1 if k == "slab_overlap" else 0)
    return s


M = 1 << 20
run("default")
run("overlap off", slab_overlap=0)
run("per-kernel timing events on", timing=1)
run("default again")
for slab, first in ((148, 74), (148, 37), (128, 64), (111, 37), (96, 32), (74, 37), (74, 74), (64, 32), (48, 24), (37, 37), (32, 16)):
    run("slab %d first %d" % (slab, first), host_slab_bytes=slab * M, host_first_slab_bytes=first * M)
for unit in (256, 512):
    run("unit %d KiB (every slab)" % unit, chunk_bytes=unit << 10)
