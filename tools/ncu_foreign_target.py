"""Small fixed workload for ncu captures of the entry-point path: two inflates of a libz level-6 stream of text-like data.
usage: ncu_foreign_target.py [mib=256]"""
import os
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import zlibk_b200
from zlibk_b200 import corpus

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench import libz_stream  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = mib << 20
data = corpus.text_like(n).tobytes()
stream = libz_stream(data, 6)
dev = torch.device("cuda", 0)
eng = zlibk_b200.Engine(0)
d_in = torch.from_numpy(np.frombuffer(stream, dtype=np.uint8).copy()).to(dev)
d_out = torch.empty(n + 64, dtype=torch.uint8, device=dev)
d_res = torch.zeros(4, dtype=torch.int64, device=dev)
for _ in range(2):
    eng.inflate_device(d_in.data_ptr(), len(stream), d_out.data_ptr(), n, d_res.data_ptr())
    eng.synchronize()
r = d_res.cpu().tolist()
print("inflate status", r[2], "bytes", r[0], "ok", bytes(d_out[:n].cpu().numpy()) == data)
assert r[2] == 1
