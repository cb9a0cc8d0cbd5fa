"""Per-role cycle breakdown of lz77_kernel (debug counters), for one corpus kind at a time."""
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
n = 148 << 20
level = int(sys.argv[1]) if len(sys.argv) > 1 else 6
strategy = int(sys.argv[2]) if len(sys.argv) > 2 else 0
only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
for name, gen in kinds.items():
    if only and name not in only:
        continue
    host = gen(n)
    d_in = torch.from_numpy(host).to(dev)
    cap = eng.compress_bound(n)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(4, dtype=torch.int64, device=dev)
    units = 256
    prof = torch.zeros(units * 32, dtype=torch.int64, device=dev)
    eng.set_option("chunk_bytes", 1 << 20)
    eng.set_option("lz_profile_buffer", prof.data_ptr())
    eng.set_option("timing", 1)
    for knob in ("lz_chain", "lz_max_lazy", "lz_nice"):
        if os.environ.get(knob.upper()):
            eng.set_option(knob, int(os.environ[knob.upper()]))
    for it in range(2):
        eng.counter("reset_timing")
        eng.deflate_device(d_in.data_ptr(), n, level, strategy, d_out.data_ptr(), cap, d_res.data_ptr())
        eng.synchronize()
    p = prof.cpu().numpy().reshape(units, 32)[: n >> 20].astype(np.float64)
    tot = p[:, 0].mean()
    lz_ms = eng.counter("lz77_us") / 1e3
    print("%-9s L%d lz77 %.2f ms (%.1f GB/s) huff %.2f pack %.2f ratio %.4f | cycles/byte %.2f | builder role %.0f%% search %.0f%% | parser role %.0f%% search %.0f%% | searcher search %.0f%%" % (
        name, level, lz_ms, n / lz_ms / 1e6, eng.counter("huffman_us") / 1e3, eng.counter("pack_us") / 1e3, d_res.cpu().tolist()[0] / n, tot / (1 << 20),
        100 * p[:, 1].mean() / tot, 100 * p[:, 2].mean() / tot, 100 * p[:, 3].mean() / tot, 100 * p[:, 4].mean() / tot, 100 * p[:, 6].mean() / tot))
    print("          parser phases %% of unit: pass A + stitch 1 %.1f wait %.1f stitch 2 %.1f pass B + count %.1f wait %.1f copy-out %.1f" % tuple(
        100 * p[:, 8 + q].mean() / tot for q in range(6)))
eng.set_option("lz_profile_buffer", 0)
