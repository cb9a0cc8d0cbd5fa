"""Per-corpus-kind timing of the inflate kernels (decode vs the rest), 148 MiB each, level 6 streams of this engine."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import zlibk_b200
from zlibk_b200 import corpus

eng = zlibk_b200.Engine(0)
dev = torch.device("cuda", 0)
kinds = {"text": lambda n: corpus.text_like(n), "random": lambda n: corpus._random_bytes(n, 5), "runs": lambda n: corpus._segment(3, n, 7),
         "repeat4k": lambda n: corpus._segment(2, n, 8), "counters": lambda n: corpus._segment(4, n, 9), "base64": lambda n: corpus._segment(5, n, 10),
         "mixed": lambda n: corpus.mixed_entropy(n)}
n = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 148 << 20
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
block = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if block:
    eng.set_option("block_bytes", block)
eng.set_option("timing", 1)
for name, gen in kinds.items():
    if only and name not in only:
        continue
    host = gen(n)
    d_in = torch.from_numpy(host).to(dev)
    cap = eng.compress_bound(n)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(4, dtype=torch.int64, device=dev)
    eng.deflate_device(d_in.data_ptr(), n, 6, 0, d_out.data_ptr(), cap, d_res.data_ptr())
    eng.synchronize()
    m = d_res.cpu().tolist()[0]
    d_back = torch.empty(n, dtype=torch.uint8, device=dev)
    d_res2 = torch.zeros(4, dtype=torch.int64, device=dev)
    for it in range(3):
        eng.counter("reset_timing")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.inflate_device(d_out.data_ptr(), m, d_back.data_ptr(), n, d_res2.data_ptr())
        e1.record()
        eng.synchronize()
        torch.cuda.synchronize()
    ok = bool(torch.equal(d_back, d_in))
    print("%-9s ratio %.4f inflate %.2f ms (%.1f GB/s) decode %.2f other %.2f ok %s" % (
        name, m / n, e0.elapsed_time(e1), n / e0.elapsed_time(e1) / 1e6, eng.counter("inflate_decode_us") / 1e3,
        eng.counter("inflate_other_us") / 1e3, ok))
