"""What the box allows end to end: N ranks, each copying 1 GiB up and 0.42 GiB down between pinned host memory and its
GPU, all at once, one cudaMemcpyAsync per 64 MiB slab -- the denominator for bench.py's end-to-end figure.
  python tools/host_ceiling.py                                   (one GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/host_ceiling.py
Optional: --bind pins every rank's threads to the CPUs nvidia-smi topo reports for its GPU (os.sched_setaffinity)."""
import json, os, subprocess, sys, time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
bind = "--bind" in sys.argv
cpus = None
if bind:
    try:
        q = subprocess.run(["nvidia-smi", "topo", "-C", "-i", str(local)], capture_output=True, text=True).stdout
        spec = q.strip().split(":")[-1].strip()
        ids = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, ids)
        cpus = sorted(ids)
    except Exception as e:
        cpus = "failed: %s" % e
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
up, down, slab = 1 << 30, int(0.42 * (1 << 30)), 64 << 20
h_up = torch.empty(up, dtype=torch.uint8).pin_memory()
h_down = torch.empty(down, dtype=torch.uint8).pin_memory()
d_up = torch.empty(slab, dtype=torch.uint8, device=dev)
d_down = torch.empty(down, dtype=torch.uint8, device=dev)
s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def copies(do_up=True, do_down=True):
    if do_up:
        with torch.cuda.stream(s_up):
            for off in range(0, up, slab):
                d_up[: min(slab, up - off)].copy_(h_up[off: off + slab], non_blocking=True)
    if do_down:
        with torch.cuda.stream(s_down):
            for off in range(0, down, slab):
                h_down[off: off + slab].copy_(d_down[off: off + slab], non_blocking=True)


def timed(**kw):
    best = 1e9
    for it in range(4):
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        copies(**kw)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        if it:
            best = min(best, dt)
    return best


both, only_up, only_down = timed(), timed(do_down=False), timed(do_up=False)
if rank == 0:
    print(json.dumps({"n_gpus": world, "bind": cpus, "per_gpu": {"h2d_bytes": up, "d2h_bytes": down},
                      "both_ms": both * 1e3, "uncompressed_gbs_ceiling": up * world / both / 1e9,
                      "h2d_alone_gbs_per_gpu": up / only_up / 1e9, "d2h_alone_gbs_per_gpu": down / only_down / 1e9,
                      "h2d_alone_gbs_total": up * world / only_up / 1e9, "d2h_alone_gbs_total": down * world / only_down / 1e9}))
if world > 1:
    dist.destroy_process_group()
