"""End-to-end host-buffer inflate of a 1 GiB (or argv[1] MiB) mixed-corpus stream: one pass against the pipeline at several
slab sizes.  ZK_PIPE_DEBUG=1 prints the per-slab timeline of the pipeline."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import zlibk_b200
from zlibk_b200 import corpus

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = mib << 20
eng = zlibk_b200.Engine(0)
host = corpus.mixed_entropy(n)
h_in = torch.from_numpy(host).pin_memory()
h_out = torch.empty(eng.compress_bound(n), dtype=torch.uint8).pin_memory()
rc, s, _ = eng.deflate(h_in, 6, out=h_out)
assert rc == 0
h_stream = torch.from_numpy(np.array(s, copy=True)).pin_memory()
h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
for pipe, slab in ((0, 0), (1, 0), (1, 24 << 20), (1, 64 << 20), (1, 96 << 20), (1, 128 << 20)):
    eng.set_option("inflate_pipe", pipe)
    eng.set_option("inflate_slab_bytes", slab)
    best = 1e9
    for it in range(4):
        t0 = time.perf_counter()
        rc, out, used = eng.inflate(h_stream, n, out=h_back)
        dt = time.perf_counter() - t0
        if it:
            best = min(best, dt)
    ok = rc == 1 and bool(torch.equal(h_back, h_in))
    print("pipe %d slab %3d MiB: %.2f ms  %.1f GB/s  rc %d bit-exact %s slabs %d fallbacks %d" % (
        pipe, slab >> 20, best * 1e3, n / best / 1e9, rc, ok, eng.counter("inflate_pipe_slabs"), eng.counter("inflate_pipe_fallbacks")), flush=True)
