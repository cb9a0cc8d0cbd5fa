"""Multi-GPU correctness check (run under torchrun): every rank deflates its shard of ONE logical input, the
bodies are gathered on rank 0 only to VERIFY that header + bodies + tail is a valid zlib stream of the whole input."""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import zlibk_b200
from zlibk_b200 import corpus, sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 96 << 20
whole = corpus.mixed_entropy(n)  # every rank can generate the input; only its shard goes to its GPU
a, b = sharding.plan_shards(n, world)[rank]
eng = zlibk_b200.Engine(local)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st); eng.set_stream(st.cuda_stream)
d_in = torch.from_numpy(whole[a:b]).to(dev)
hist = torch.from_numpy(whole[max(0, a - 32768):a].copy()).to(dev) if a else None
cap = eng.compress_bound(b - a)
d_body = torch.empty(cap, dtype=torch.uint8, device=dev)
sh = sharding.ShardedDeflate(eng, rank, world, dev, hist)
assert sh.step(d_in, b - a, 6, d_body, cap) == 0
torch.cuda.synchronize(dev)
offsets, total, adler = sh.finish()
assert sh.last_status == 0
# verification only: collect the bodies on rank 0 (the product leaves them sharded)
lens = [int(x) for x in sh.gathered[:, 0].cpu().tolist()]
mx = max(lens)
pad = torch.zeros(mx, dtype=torch.uint8, device=dev); pad[: lens[rank]] = d_body[: lens[rank]]
bufs = [torch.empty(mx, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
dist.gather(pad, bufs, dst=0)
if rank == 0:
    stream = eng.zlib_header(6) + b"".join(bufs[r][: lens[r]].cpu().numpy().tobytes() for r in range(world)) + eng.zlib_tail(adler)
    assert len(stream) == total and offsets[0] == 2
    assert adler == zlib.adler32(whole.tobytes())
    assert zlib.decompress(stream) == whole.tobytes()
    one = zlibk_b200.Engine(local)
    rc, back, used = one.inflate(stream, n)
    assert rc == 1 and bytes(back) == whole.tobytes()
    print("mgpu ok: world %d, %d -> %d bytes (%.4f), offsets %s" % (world, n, total, total / n, offsets))
# sharded inflate of the same stream: every rank decodes its body, one all_gather of tail maps, every rank gets its range
d_back = torch.empty(b - a, dtype=torch.uint8, device=dev)
si = sharding.ShardedInflate(eng, rank, world, dev)
assert si.step(d_body, lens[rank], d_back, b - a) == 0
torch.cuda.synchronize(dev)
got, status, a32, metas = si.finish()
assert status == 1 and got == b - a and all(st == 0 for _, st in metas), (rank, status, got, metas)
assert torch.equal(d_back, d_in)
parts = torch.tensor([a32, b - a], dtype=torch.int64, device=dev)
allp = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
dist.all_gather(allp, parts)
folded = 1
for t in allp:
    x = t.cpu().tolist()
    folded = sharding.adler32_combine(folded, x[0] & 0xFFFFFFFF, x[1])
assert folded == adler
if rank == 0:
    print("mgpu inflate ok: world %d, every rank reproduced its range, folded Adler-32 %08x" % (world, folded))
dist.barrier()
dist.destroy_process_group()
