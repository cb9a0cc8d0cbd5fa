#!/usr/bin/env python
"""bench.py -- deflate / inflate throughput of the B200 engine on BASELINE.json's workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size-mib M] [--level L]

One "step" = one level-6 zlib deflate of the whole synthetic buffer (the hot path), device-resident:
the input is already in HBM when the timed region starts and the stream stays in HBM.  N = 1 uses
BASELINE.json configs[1] (1 GiB mixed-entropy, level 6); N > 1 shards chunk ranges across ranks
(weak scaling: 1 GiB per rank), with one all_gather of {body_len, adler, raw_len} per step.
Prints ONE JSON line on rank 0.  `value` = uncompressed GB/s of deflate (10^9 B/s, whole job);
`inflate` carries the same measurement for the inflate of the produced stream; `e2e` is deflate
through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "deflate uncompressed GB/s at level 6 (zlib-wrapped, chunked)"
UNIT = "GB/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size-mib", type=int, default=1024, help="uncompressed MiB per GPU")
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--workload", default="mixed", choices=["mixed", "text"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --size-mib per GPU (the driver's default); strong: --total-gib cut across the GPUs (BASELINE.json configs[3])")
    ap.add_argument("--total-gib", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-mib", type=int, default=32)
    ap.add_argument("--config1-mib", type=int, default=64, help="size of the configs[0] CPU round trip inside cpu_baseline (0 = skip)")
    ap.add_argument("--no-verify", action="store_true", help="skip handing the timed region's stream to libz")
    ap.add_argument("--foreign-mib", type=int, default=256,
                    help="inflate leg on a stream libz made of the first MiB of the corpus (no sync markers); 0 = skip")
    return ap.parse_args()


PIECE = 1 << 30  # strong scaling: the job is a sequence of 1 GiB pieces, piece j made with the seed of weak-scaling rank j


def make_corpus(kind, nbytes, rank=0, first_piece=None):
    from zlibk_b200 import corpus

    if first_piece is not None:  # strong scaling: pieces first_piece .. of the global buffer
        import numpy as np

        parts = [make_corpus(kind, min(PIECE, nbytes - off), first_piece + off // PIECE) for off in range(0, nbytes, PIECE)]
        return parts[0] if len(parts) == 1 else np.concatenate(parts)
    seed_shift = 0x10000 * rank
    if kind == "text":
        return corpus.text_like(nbytes, corpus.SEED_TEXT + seed_shift)
    return corpus.mixed_entropy(nbytes, corpus.SEED_MIXED + seed_shift)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()  # the exact process we started
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _oracle():
    from tests.oracle_binding import Oracle

    so = os.path.join(ROOT, "oracle", "libzk_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return Oracle(so)


def libz_pigz_style(data, level, threads, chunk=128 << 10):
    """B3 of SURVEY.md 8(d): what pigz does, with libz -- independent chunks, each primed with the previous 32 KiB as its
    dictionary and ended with a sync flush, compressed on `threads` host threads (libz releases the GIL).  Returns
    (compressed bytes without the zlib wrapper, seconds)."""
    import zlib
    from concurrent.futures import ThreadPoolExecutor

    mv = memoryview(data)

    def one(off):
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, zlib.Z_DEFAULT_STRATEGY, bytes(mv[max(0, off - 32768): off])) if off \
            else zlib.compressobj(level, zlib.DEFLATED, -15)
        return c.compress(mv[off: off + chunk]) + c.flush(zlib.Z_SYNC_FLUSH)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(one, range(0, len(data), chunk)))
    return sum(len(p) for p in parts), time.perf_counter() - t0


def cpu_baseline(kind, level, sample_mib, config1_mib=64):
    """CPU numbers beside the GPU ones, on bounded samples of the same workload:
      B1  the oracle (C restatement of the reference algorithm, kind "port"), one host core -- the reference is
          single-threaded and could not be built here; with what the reference's own inflate makes of that stream;
      B2  libz 1.3 at the same level, one thread;  B3  libz pigz-style on every host core;
      C1  BASELINE.json configs[0]: the reference path on a 64 MiB text-like buffer at level 6, round trip."""
    import zlib

    oracle = _oracle()
    cores = os.cpu_count() or 1
    data = make_corpus(kind, sample_mib << 20).tobytes()
    t0 = time.perf_counter()
    rc, stream, used = oracle.deflate(data, level)
    dt = time.perf_counter() - t0
    rc_i, out_i, second_i, _ = oracle.inflate(stream, cap=len(data) + 64)
    try:
        libz_ok = zlib.decompress(stream) == data
        libz_says = "ok" if libz_ok else "wrong bytes"
    except zlib.error as e:
        libz_says = str(e)
    t0 = time.perf_counter()
    z = zlib.compress(data, level)
    dz = time.perf_counter() - t0
    t0 = time.perf_counter()
    zlib.decompress(z)
    dzi = time.perf_counter() - t0
    pz_len, dpz = libz_pigz_style(data, level, cores)
    out = {
        "value": len(data) / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
        "sample": "first %d MiB of the %s corpus, level %d, one pass" % (sample_mib, kind, level),
        "deflate_rc": rc, "bytes_in": used,
        "reference_stream_ratio": len(stream) / len(data),
        "reference_round_trip": {"rc": rc_i, "bytes_out": len(out_i), "bit_exact": rc_i == 1 and out_i == data, "libz_on_that_stream": libz_says,
                                 "note": "the reference's own inflate (oracle) on the reference's own stream; rc -3 with 0 bytes means the "
                                         "stream is invalid from its first block (stale-window compare, DS:672-691), so its size is "
                                         "the size of an invalid stream"},
        "libz_1thread": {"value": len(data) / dz / 1e9, "unit": UNIT, "cores": 1, "ratio": len(z) / len(data),
                         "inflate_gbs": len(data) / dzi / 1e9, "kind": "libz 1.3 level %d" % level},
        "libz_pigz_style": {"value": len(data) / dpz / 1e9, "unit": UNIT, "cores": cores, "ratio": pz_len / len(data),
                            "kind": "libz 1.3 level %d, 128 KiB chunks, 32 KiB dictionary + sync flush per chunk, one thread per host core" % level},
        "note": "value = B1, the C restatement of the reference algorithm (single thread, -O2): an upper bound on the Kotlin/Native "
                "binary, which could not be built in this image; libz_1thread / libz_pigz_style are the industry CPU baselines",
    }
    if config1_mib:
        text = make_corpus("text", config1_mib << 20).tobytes()
        t0 = time.perf_counter()
        rc1, s1, used1 = oracle.deflate(text, 6)
        d1 = time.perf_counter() - t0
        t0 = time.perf_counter()
        rc1i, out1, _, _ = oracle.inflate(s1, cap=len(text) + 64)
        d1i = time.perf_counter() - t0
        t0 = time.perf_counter()
        z1 = zlib.compress(text, 6)
        dz1 = time.perf_counter() - t0
        t0 = time.perf_counter()
        back = zlib.decompress(z1)
        dz1i = time.perf_counter() - t0
        out["config1"] = {
            "workload": "%d MiB text-like buffer (seed 0x5EED0001), level 6, compress + decompress on the CPU (BASELINE.json configs[0])" % config1_mib,
            "reference": {"deflate_rc": rc1, "bytes_in": used1, "deflate_gbs": len(text) / d1 / 1e9, "stream_ratio": len(s1) / len(text),
                          "inflate_rc": rc1i, "inflate_bytes_out": len(out1), "inflate_seconds": d1i,
                          "round_trip_ok": rc1i == 1 and out1 == text, "kind": "port", "cores": 1},
            "libz": {"deflate_gbs": len(text) / dz1 / 1e9, "inflate_gbs": len(text) / dz1i / 1e9, "ratio": len(z1) / len(text),
                     "round_trip_ok": back == text, "cores": 1},
        }
    return out


def libz_stream(data, level, piece=32 << 20):
    """ONE zlib stream of `data` whose blocks libz made: pieces of 32 MiB are deflated on all host threads as raw bodies primed
    with the 32 KiB before them and joined (a sync flush per join: 8 markers in 256 MiB, nothing an inflater could use)."""
    import zlib
    from concurrent.futures import ThreadPoolExecutor

    view = memoryview(data)

    def one(i):
        lo = i * piece
        hi = min(len(data), lo + piece)
        if lo:
            co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, zlib.Z_DEFAULT_STRATEGY, bytes(view[lo - 32768: lo]))
        else:
            co = zlib.compressobj(level, zlib.DEFLATED, -15)
        return co.compress(view[lo:hi]) + co.flush(zlib.Z_FINISH if hi == len(data) else zlib.Z_SYNC_FLUSH)

    with ThreadPoolExecutor(os.cpu_count() or 8) as ex:
        bodies = list(ex.map(one, range((len(data) + piece - 1) // piece)))
    return b"\x78\x9c" + b"".join(bodies) + zlib.adler32(data).to_bytes(4, "big")


def foreign_leg(eng, dev, host, nbytes, level):
    """Inflate of a stream this engine did not make: the first `nbytes` of the corpus deflated by libz.  Such a stream has no
    entry points; the engine finds the dynamic block headers itself (zk_inflate.cu block_find_kernel)."""
    import numpy as np
    import torch

    data = host[:nbytes].tobytes()
    stream = libz_stream(data, level)
    m = len(stream)
    d_s = torch.from_numpy(np.frombuffer(stream, dtype=np.uint8).copy()).to(dev)
    d_b = torch.empty(nbytes + 64, dtype=torch.uint8, device=dev)
    d_r = torch.zeros(4, dtype=torch.int64, device=dev)
    eng.set_option("timing", 1)
    c0, m0 = eng.counter("inflate_foreign_chunks"), eng.counter("inflate_foreign_merges")
    best = None
    for i in range(4):
        torch.cuda.synchronize(dev)
        eng.counter("reset_timing")
        t0 = time.perf_counter()
        eng.inflate_device(d_s.data_ptr(), m, d_b.data_ptr(), nbytes, d_r.data_ptr())
        eng.synchronize()
        dt = time.perf_counter() - t0
        if i and (best is None or dt < best[0]):
            best = (dt, eng.counter("inflate_decode_us") / 1e3, eng.counter("inflate_other_us") / 1e3)
    eng.set_option("timing", 0)
    r = d_r.cpu().tolist()
    ok = r[2] == 1 and r[0] == nbytes and bytes(d_b[:nbytes].cpu().numpy()) == data
    runs = 4
    h_s = torch.from_numpy(np.frombuffer(stream, dtype=np.uint8).copy()).pin_memory()
    h_b = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    eng.inflate(h_s, nbytes, out=h_b)
    t0 = time.perf_counter()
    rc, _, _ = eng.inflate(h_s, nbytes, out=h_b)
    e2e = time.perf_counter() - t0
    eng.set_option("inflate_foreign", 0)
    small = min(nbytes, 16 << 20)  # the one-warp decoder the engine had before, on a stream it can finish in a second
    s_small = libz_stream(data[:small], level, piece=small)
    t0 = time.perf_counter()
    rc0, _, _ = eng.inflate(s_small, small)
    t_one_warp = time.perf_counter() - t0
    eng.set_option("inflate_foreign", 1)
    return {"stream": "libz 1.3 level %d of the first %d MiB of the corpus: no sync markers" % (level, nbytes >> 20), "compressed_bytes": m,
            "value": nbytes / best[0] / 1e9, "unit": UNIT, "ms": best[0] * 1e3, "decode_kernel_ms": best[1], "other_kernels_ms": best[2],
            "entry_points": (eng.counter("inflate_foreign_chunks") - c0) // (runs + 2) - 1,
            "false_entry_points_merged": (eng.counter("inflate_foreign_merges") - m0) // (runs + 2), "bit_exact": bool(ok), "status": r[2],
            "e2e": {"value": nbytes / e2e / 1e9, "unit": UNIT, "status": rc},
            "without_entry_points": {"value": small / t_one_warp / 1e9, "unit": UNIT, "bytes": small, "status": rc0,
                                     "note": "option inflate_foreign = 0: the whole stream is one chunk for one warp"}}


def workload_config(args, world):
    """The `config` object both arms print: the workload the metric is quoted on."""
    kind = "mixed-entropy" if args.workload == "mixed" else "text-like"
    seed = 2 if args.workload == "mixed" else 1
    if args.scaling == "strong":
        return {"workload": "%d GiB %s corpus in all (seed 0x5EED000%d, 1 GiB pieces), level %d, ONE zlib stream cut into chunk ranges, 32 KiB blocks; "
                            "input larger than L2 (no flush needed)" % (args.total_gib, kind, seed, args.level),
                "parallelism": "chunk ranges per GPU x%d (strong scaling: %d MiB per GPU)" % (world, args.total_gib * 1024 // world)}
    return {"workload": "%d MiB %s corpus per GPU (seed 0x5EED000%d), level %d, zlib stream, 32 KiB blocks; input larger than L2 (no flush needed)" % (
                args.size_mib, kind, seed, args.level),
            "parallelism": "chunk ranges per GPU x%d" % world}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Kotlin reference can
    not be compiled here), one host thread -- the reference is single-threaded -- on a bounded sample per step."""
    if rank != 0:
        return
    oracle = _oracle()
    sample_mib = min(args.cpu_sample_mib, 16)
    data = make_corpus(args.workload, sample_mib << 20).tobytes()
    for _ in range(min(args.warmup, 1)):
        oracle.deflate(data[: 1 << 20], args.level)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        oracle.deflate(data, args.level)
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    value = len(data) / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "each step deflates the first %d MiB of the %s corpus on one host core (the reference is "
                                   "single-threaded: one core is every thread it can use)" % (sample_mib, args.workload)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import zlibk_b200
    from zlibk_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    strong = args.scaling == "strong"
    if strong:
        if (args.total_gib * 1024) % world:
            raise SystemExit("--total-gib must divide by the number of GPUs")
        n = (args.total_gib << 30) // world
        if n % PIECE and PIECE % n:
            raise SystemExit("strong scaling: the per-GPU share must be a multiple or a divisor of 1 GiB")
        if n >= PIECE:
            host = make_corpus(args.workload, n, 0, first_piece=rank * (n // PIECE))
        else:  # several ranks share a piece: every rank builds the piece and keeps its slice
            per = PIECE // n
            host = make_corpus(args.workload, PIECE, rank // per)[(rank % per) * n: (rank % per + 1) * n].copy()
    else:
        n = args.size_mib << 20
        host = make_corpus(args.workload, n, rank)
    h_in = torch.from_numpy(host).pin_memory()
    d_in = h_in.to(dev, non_blocking=True)
    eng = zlibk_b200.Engine(local_rank)
    # a dedicated stream: torch's default stream has handle 0, which the C ABI reads as "use the private stream"
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    cap = eng.compress_bound(n)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(4, dtype=torch.int64, device=dev)
    d_back = torch.empty(n, dtype=torch.uint8, device=dev)
    d_res2 = torch.zeros(4, dtype=torch.int64, device=dev)
    torch.cuda.synchronize(dev)

    shard = None
    if world > 1:
        # the 32 KiB halo comes from the source (here: the previous rank's generator), not from a collective
        hist = None
        if rank > 0:
            from zlibk_b200 import corpus
            if not strong:
                hist = torch.from_numpy(corpus.tail(args.workload, n, rank - 1, 32768)).to(dev)
            elif n >= PIECE:
                hist = torch.from_numpy(corpus.tail(args.workload, PIECE, rank * (n // PIECE) - 1, 32768)).to(dev)
            else:
                per = PIECE // n
                piece = make_corpus(args.workload, PIECE, (rank - 1) // per)
                hist = torch.from_numpy(piece[((rank - 1) % per + 1) * n - 32768: ((rank - 1) % per + 1) * n].copy()).to(dev)
                del piece
        shard = sharding.ShardedDeflate(eng, rank, world, dev, hist)

    def deflate_step():
        if shard is None:
            rc = eng.deflate_device(d_in.data_ptr(), n, args.level, 0, d_out.data_ptr(), cap, d_res.data_ptr())
        else:
            rc = shard.step(d_in, n, args.level, d_out, cap)
        if rc != 0:
            raise SystemExit("deflate failed rc=%d %s" % (rc, eng.last_error()))

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        deflate_step()
    barrier()
    eng.counter("reset_timing")
    launches0 = eng.counter("kernel_launches")
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        deflate_step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    launches = eng.counter("kernel_launches") - launches0
    # ---- per-kernel durations: the same steps once more with CUDA events around every launch (on the stream it is
    # launched on) and the slabs of a step strictly one after the other.  In the timed region above the kernels of two
    # slabs overlap on two streams, where the span of a single kernel is not separable from its neighbours'.
    eng.set_option("timing", 1)
    eng.set_option("slab_overlap", 0)
    deflate_step()
    barrier()
    eng.counter("reset_timing")
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(args.steps):
        deflate_step()
    p1.record(stream)
    barrier()
    serial_ms_per_step = p0.elapsed_time(p1) / args.steps
    eng.set_option("slab_overlap", 1)
    lz_us, lz_spans = eng.counter("lz77_us"), eng.counter("lz77_spans")
    huff_us, pack_us, adler_us = eng.counter("huffman_us"), eng.counter("pack_us"), eng.counter("adler_us")
    deflate_step()  # leave the overlapped run's stream in d_out for the legs below
    barrier()
    if shard is None:
        res = d_res.cpu().tolist()
        out_len, status = res[0], res[2]
    else:
        offsets, stream_len, folded = shard.finish()
        out_len, status = shard.last_body_len, shard.last_status
    if status != 0:
        raise SystemExit("deflate status %d" % status)
    # the stream the timed region produced goes to libz once (the whole stream, on the host): nothing but the engine's own
    # inflate would otherwise have seen it
    verify = None
    if shard is None and not args.no_verify:
        import zlib

        t0 = time.perf_counter()
        back = zlib.decompress(d_out[:out_len].cpu().numpy().tobytes())
        verify = {"libz_inflate_of_the_timed_stream": "bit-exact" if back == host.tobytes() else "MISMATCH",
                  "bytes": len(back), "seconds": time.perf_counter() - t0}
        del back
        if verify["libz_inflate_of_the_timed_stream"] != "bit-exact":
            raise SystemExit("libz does not reproduce the input from the timed region's stream")
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tot = torch.tensor([float(out_len)], dtype=torch.float64, device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        total_out = float(tot.item())
    else:
        total_out = float(out_len)
    ms_per_step = ms / args.steps
    value = (n * world) / (ms_per_step * 1e-3) / 1e9

    # ---- inflate (device-resident).  N = 1: the stream just produced.  N > 1: the SAME sharded stream -- every rank holds
    # the body of its chunk range and ends up with the bytes of that range; the only collective is one all_gather of the
    # 64 KiB tail maps (zlibk_b200.sharding.ShardedInflate).
    inflate = None
    isteps = max(1, min(args.steps, 3))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if shard is not None:
        m = int(shard.last_body_len)
        sinf = sharding.ShardedInflate(eng, rank, world, dev)

        def inflate_step():
            rc = sinf.step(d_out, m, d_back, n)
            if rc != 0:
                raise SystemExit("sharded inflate failed rc=%d %s" % (rc, eng.last_error()))
    else:
        m = int(out_len)

        def inflate_step():
            eng.inflate_device(d_out.data_ptr(), m, d_back.data_ptr(), n, d_res2.data_ptr())
    for _ in range(min(args.warmup, 2)):
        inflate_step()
    barrier()
    eng.counter("reset_timing")
    e0.record(stream)
    for _ in range(isteps):
        inflate_step()
    e1.record(stream)
    barrier()
    ims = e0.elapsed_time(e1) / isteps
    if shard is not None:
        got_len, istatus, _, _ = sinf.finish()
        ok = istatus == 1 and got_len == n and bool(torch.equal(d_back, d_in))
    else:
        r2 = d_res2.cpu().tolist()
        istatus = r2[2]
        ok = r2[2] == 1 and r2[0] == n and bool(torch.equal(d_back, d_in))
    if world > 1:
        t = torch.tensor([ims, 0.0 if ok else 1.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ims, ok = float(t[0].item()), float(t[1].item()) == 0.0
    inflate = {"value": n * world / (ims * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ims, "status": istatus, "bit_exact": ok,
               "decode_kernel_ms": eng.counter("inflate_decode_us") / 1e3 / isteps,
               "other_kernels_ms": eng.counter("inflate_other_us") / 1e3 / isteps,
               "streams": "one stream" if world == 1 else "one stream across %d GPUs (chunk ranges; one all_gather of 64 KiB tail maps)" % world}
    if shard is not None:  # the host-buffer leg below inflates a complete stream of this rank's range
        if eng.deflate_device(d_in.data_ptr(), n, args.level, 0, d_out.data_ptr(), cap, d_res.data_ptr()) != 0:
            raise SystemExit("deflate for the inflate end-to-end leg failed: %s" % eng.last_error())
        eng.synchronize()
        m = int(d_res.cpu().tolist()[0])
    eng.set_option("timing", 0)  # the end-to-end legs run the calls as a user gets them: no profiling events
    # end to end through the host-buffer call (pinned memory, H2D of the stream and D2H of the output inside)
    h_stream = d_out[:m].cpu().pin_memory()
    h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
    eng.inflate(h_stream, n, out=h_back)  # warm-up: staging allocation
    barrier()
    t0 = time.perf_counter()
    rc_i, back_i, _ = eng.inflate(h_stream, n, out=h_back)
    torch.cuda.synchronize(dev)
    idt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([idt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        idt = float(t.item())
    inflate["e2e"] = {"value": n * world / idt / 1e9, "unit": UNIT, "h2d_bytes_per_step": m, "d2h_bytes_per_step": n, "status": rc_i}
    del h_stream, h_back
    if world == 1 and args.foreign_mib > 0:
        inflate["libz_made_stream"] = foreign_leg(eng, dev, host, min(n, args.foreign_mib << 20), args.level)

    # ---- e2e through the host-buffer C-ABI call (pinned memory; H2D + D2H inside) ----------------------------
    e2e = None
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    e_steps = max(1, min(args.steps, 3))
    h_hist = hist.cpu().pin_memory() if (world > 1 and rank > 0) else None

    def e2e_step():
        if shard is None:
            rc, s, _ = eng.deflate(h_in, args.level, out=h_out)
        else:
            rc = shard.step_host(h_in, args.level, h_out, h_hist)
            s = None
        if rc != 0:
            raise SystemExit("host deflate failed rc=%d %s" % (rc, eng.last_error()))
        return s

    e2e_step()  # warm-up: staging allocation
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        s = e2e_step()
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / e_steps
    e2e_out = int(len(s)) if s is not None else int(shard.d_meta.cpu().tolist()[0])
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": n * world / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": n, "d2h_bytes_per_step": e2e_out + 32,
           "ms_per_step": dt * 1e3, "note": "per GPU bytes; wall clock around the host-buffer call (max over ranks)"}
    # what the box allows for these bytes: the same H2D and D2H volumes as bare copies from pinned memory, every rank at
    # once, nothing else running (the denominator for the end-to-end figure)
    s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    d_tmp = torch.empty(min(n, 64 << 20), dtype=torch.uint8, device=dev)

    def bare_copies():
        with torch.cuda.stream(s_up):
            for off in range(0, n, d_tmp.numel()):
                k = min(d_tmp.numel(), n - off)
                d_tmp[:k].copy_(h_in[off:off + k], non_blocking=True)
        with torch.cuda.stream(s_down):
            for off in range(0, e2e_out, 64 << 20):
                k = min(64 << 20, e2e_out - off)
                h_out[off:off + k].copy_(d_out[off:off + k], non_blocking=True)

    bare_copies()
    barrier()
    t0 = time.perf_counter()
    bare_copies()
    torch.cuda.synchronize(dev)
    cdt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([cdt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cdt = float(t.item())
    e2e["copy_ceiling"] = {"value": n * world / cdt / 1e9, "unit": UNIT, "ms": cdt * 1e3,
                           "what": "bare cudaMemcpyAsync of the same bytes (H2D input, D2H stream) from pinned memory, both directions at once, all ranks at once"}
    e2e["frac_of_copy_ceiling"] = (n * world / dt) / (n * world / cdt)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # dominant kernel: lz77_kernel.  Algorithmic bytes of the launches timed = uncompressed + compressed bytes of
    # this rank's steps; duration = CUDA-event time around those launches on their stream.
    traffic, traffic_source = None, None
    try:  # DRAM bytes of one lz77_kernel launch from the committed ncu capture, scaled to this run's launch size
        tr = json.load(open(os.path.join(ROOT, "profiles", "lz77_traffic.json")))
        launch_bytes = n * args.steps / max(1, lz_spans)
        traffic = tr["dram_bytes_per_input_byte"] * launch_bytes
        traffic_source = tr["source"]
    except Exception:
        pass
    alg_bytes = (n + total_out / world) * args.steps
    lz_s = lz_us * 1e-6
    achieved = alg_bytes / lz_s / 1e9 if lz_s > 0 else None
    roofline = {"bound": "hbm", "kernel": "lz77_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if achieved else None, "peak_source": peak_kind, "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_source": traffic_source,
                "algorithmic_bytes_per_launch": alg_bytes / max(1, lz_spans),
                "launches_timed": lz_spans, "avg_launch_ms": lz_us / 1e3 / max(1, lz_spans),
                # the one kernel of the path that IS bandwidth-bound: Adler-32 over the input (N algorithmic bytes)
                "adler32_kernels": {"achieved": (n * args.steps) / (adler_us * 1e-6) / 1e9 if adler_us > 0 else None,
                                    "frac": (n * args.steps) / (adler_us * 1e-6) / 1e9 / peak if adler_us > 0 else None, "unit": "GB/s"},
                "timed_in": "instrumented pass after the timed region: same steps, CUDA events around every launch, slabs "
                            "one after the other (the timed region overlaps the kernels of two slabs on two streams)",
                "instrumented_ms_per_step": serial_ms_per_step,
                "share_of_step": {"lz77": lz_us / 1e3 / args.steps / serial_ms_per_step, "huffman+offsets": huff_us / 1e3 / args.steps / serial_ms_per_step,
                                  "pack": pack_us / 1e3 / args.steps / serial_ms_per_step, "adler": adler_us / 1e3 / args.steps / serial_ms_per_step}}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": workload_config(args, world),
        "compressed_ratio": total_out / (n * world),
        "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
        "inflate": inflate, "e2e": e2e, "verify": verify,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.workload, args.level, args.cpu_sample_mib, args.config1_mib)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
