"""Small fixed workload for ncu captures: one deflate (+ optional inflate) of a synthetic buffer on cuda:0.
usage: ncu_target.py [kind=text|mixed] [mib=148] [level=6] [inflate=0|1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import zlibk_b200
from zlibk_b200 import corpus

kind = sys.argv[1] if len(sys.argv) > 1 else "text"
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 148
level = int(sys.argv[3]) if len(sys.argv) > 3 else 6
do_inflate = int(sys.argv[4]) if len(sys.argv) > 4 else 0
n = mib << 20
seg = {"random": 1, "repeat4k": 2, "runs": 3, "counters": 4, "base64": 5}
host = corpus.text_like(n) if kind == "text" else corpus.mixed_entropy(n) if kind == "mixed" else corpus._segment(seg[kind], n, 10)
dev = torch.device("cuda", 0)
eng = zlibk_b200.Engine(0)
d_in = torch.from_numpy(host).to(dev)
cap = eng.compress_bound(n)
d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
d_res = torch.zeros(4, dtype=torch.int64, device=dev)
for _ in range(2):
    rc = eng.deflate_device(d_in.data_ptr(), n, level, 0, d_out.data_ptr(), cap, d_res.data_ptr())
    eng.synchronize()
res = d_res.cpu().tolist()
print("deflate rc", rc, "out", res[0], "status", res[2], "ratio %.4f" % (res[0] / n))
assert rc == 0 and res[2] == 0
if do_inflate:
    d_back = torch.empty(n, dtype=torch.uint8, device=dev)
    d_res2 = torch.zeros(4, dtype=torch.int64, device=dev)
    for _ in range(2):
        eng.inflate_device(d_out.data_ptr(), res[0], d_back.data_ptr(), n, d_res2.data_ptr())
        eng.synchronize()
    r2 = d_res2.cpu().tolist()
    print("inflate status", r2[2], "ok", bool(torch.equal(d_back, d_in)))
    assert r2[2] == 1
