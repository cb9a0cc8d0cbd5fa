"""Chunk-range sharding of one zlib stream across GPUs.

Every rank deflates a contiguous range of the input as a headerless, byte-aligned body primed with the
32 KiB that precede it (zk_deflate_shard_device).  The ONLY exchange is one all_gather of
{body_len, adler32(shard), status, raw_len} per rank; every rank then computes the same exclusive scan (its
body's offset in the stream) and folds the checksums with zlib's combine identity.  Rank 0 owns the
2-byte header, the last rank the final empty block + Adler-32 trailer.  No other collective is used.
"""
BASE = 65521


def adler32_combine(adler1, adler2, len2):
    """zlib's adler32_combine on three integers (same arithmetic as zk_adler32_combine)."""
    rem = len2 % BASE
    s1, s2 = adler1 & 0xFFFF, (adler1 >> 16) & 0xFFFF
    t1 = (adler2 & 0xFFFF) + BASE - 1
    t2 = ((adler2 >> 16) & 0xFFFF) + BASE - rem
    s2 = (s2 + rem * s1 + t2) % BASE
    s1 = (s1 + t1) % BASE
    return (s2 << 16) | s1


def plan_shards(n, world, align=32768):
    """Contiguous [begin, end) per rank, boundaries aligned to the deflate block span."""
    per = (n + world - 1) // world
    per = (per + align - 1) // align * align
    return [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]


def assemble(metas):
    """metas: per rank (body_len, adler, status, raw_len) -- the record zk_deflate_shard_device writes -- in rank
    order -> (offsets of each body in the stream, total stream length, adler of the whole input)."""
    offsets, at, adler = [], 2, 1
    for body_len, a, status, raw_len in metas:
        if status != 0:
            raise RuntimeError("a shard failed with zlib code %d" % status)
        offsets.append(at)
        at += int(body_len)
        adler = adler32_combine(adler, int(a) & 0xFFFFFFFF, int(raw_len))
    return offsets, at + 6, adler


def exchange(meta_tensor, world):
    """The one collective: all_gather of the 4-word meta record {body_len, adler, status, raw_len}.  meta_tensor:
    int64[4] on the rank's device (NCCL) or host (gloo).  Returns int64[world, 4]."""
    import torch
    import torch.distributed as dist

    out = torch.empty(world * 4, dtype=torch.int64, device=meta_tensor.device)
    dist.all_gather_into_tensor(out, meta_tensor.contiguous())
    return out.view(world, 4)


def _same_stream(engine, device):
    """The engine's kernels and torch's copies / collectives must be ordered: run the engine on torch's current stream of
    the device.  torch's DEFAULT stream has handle 0, which the C ABI reads as "the context's private stream" -- then the
    caller has to make a stream current first (bench.py does), and this raises rather than race."""
    import torch

    handle = torch.cuda.current_stream(device).cuda_stream
    if handle == 0:
        raise RuntimeError("make a non-default torch.cuda.Stream current before building a Sharded* driver: the engine has to "
                           "share a stream with torch's copies and collectives")
    engine.set_stream(handle)


class ShardedDeflate:
    """Per-rank driver used by bench.py and the multi-GPU tests."""

    def __init__(self, engine, rank, world, device, hist=None):
        import torch

        self.eng, self.rank, self.world, self.dev = engine, rank, world, device
        _same_stream(engine, device)
        self.d_meta = torch.zeros(4, dtype=torch.int64, device=device)
        self.d_hist = hist  # uint8 tensor with the 32 KiB before this shard, or None for rank 0
        self.gathered = None
        self.last_body_len = 0
        self.last_status = 0

    def step(self, d_in, n, level, d_body, cap):
        hist_ptr = self.d_hist.data_ptr() if self.d_hist is not None else None
        hist_len = int(self.d_hist.numel()) if self.d_hist is not None else 0
        rc = self.eng.deflate_shard_device(hist_ptr, hist_len, d_in.data_ptr(), n, level, 0, d_body.data_ptr(), cap, self.d_meta.data_ptr())
        if rc != 0:
            return rc
        self.gathered = exchange(self.d_meta, self.world)  # the only collective
        return 0

    def step_host(self, h_in, level, h_body, h_hist=None):
        """The same through host buffers (H2D + D2H inside, pipelined in slabs): the end-to-end leg of bench.py.  The
        meta record is rebuilt on the host and exchanged with the same single all_gather."""
        import torch

        rc, body, adler = self.eng.deflate_shard(h_hist, h_in, level, 0, out=h_body)
        if rc != 0:
            return rc
        meta = torch.tensor([len(body), adler, 0, int(h_in.numel() if hasattr(h_in, "numel") else len(h_in))], dtype=torch.int64)
        self.d_meta.copy_(meta)
        self.gathered = exchange(self.d_meta, self.world)
        return 0

    def finish(self):
        """Host view after the timed region: offsets, total length, folded Adler-32."""
        g = self.gathered.cpu().tolist()
        meta = self.d_meta.cpu().tolist()
        self.last_body_len, self.last_status = meta[0], meta[2]
        return assemble([tuple(row) for row in g])

    @property
    def ready(self):
        return self.gathered is not None


class ShardedInflate:
    """One stream inflated across GPUs.  Rank r holds the body of its chunk range (what ShardedDeflate made) and ends up
    with the bytes of that range.  The only collective is one all_gather of a 64 KiB + 32 B record per rank: the shard's
    tail map (see include/zlibk_gpu.h) and {out_len, adler, status, consumed}."""

    RECORD = 65536 + 32

    def __init__(self, engine, rank, world, device):
        import torch

        self.eng, self.rank, self.world, self.dev = engine, rank, world, device
        _same_stream(engine, device)
        self.d_record = torch.zeros(self.RECORD, dtype=torch.uint8, device=device)   # tail map | meta of this rank
        self.d_all = torch.zeros(world * self.RECORD, dtype=torch.uint8, device=device)
        self.d_entry = torch.zeros(32768, dtype=torch.uint8, device=device)
        self.d_res = torch.zeros(4, dtype=torch.int64, device=device)

    def step(self, d_body, m, d_out, cap):
        """-> zlib code of the C ABI calls (0 = launched); results via finish()."""
        import torch
        import torch.distributed as dist

        hist_len = 0 if self.rank == 0 else 32768
        rc = self.eng.inflate_shard_prepare_device(d_body.data_ptr(), m, hist_len, cap, self.d_record.data_ptr(), self.d_res.data_ptr())
        if rc != 0:
            return rc
        # the one collective: every rank's tail map (the meta words ride along for the host)
        self.d_record[65536:].view(torch.int64).copy_(self.d_res)
        dist.all_gather_into_tensor(self.d_all, self.d_record)
        rc = self.eng.window_from_maps_device(self.d_all.data_ptr(), self.RECORD, self.rank, self.d_entry.data_ptr())
        if rc != 0:
            return rc
        return self.eng.inflate_shard_finish_device(self.d_entry.data_ptr(), d_out.data_ptr(), self.d_res.data_ptr())

    def finish(self):
        """-> (this rank's out_len, status, adler of its output, [(out_len, status) of every rank after phase 1])."""
        import torch

        out_len, _, status, adler = self.d_res.cpu().tolist()
        metas = self.d_all.view(self.world, self.RECORD)[:, 65536:].contiguous().view(torch.int64).view(self.world, 4).cpu().tolist()
        return out_len, status, adler & 0xFFFFFFFF, [(mm[0], mm[2]) for mm in metas]
