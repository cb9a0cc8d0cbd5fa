"""CPU tests for the deterministic corpora and the host side of the multi-GPU path (world_size 2, gloo)."""
import os
import socket
import zlib

import numpy as np
import pytest


def test_corpora_are_deterministic():
    from zlibk_b200 import corpus

    a = corpus.mixed_entropy(3 << 20)
    b = corpus.mixed_entropy(3 << 20)
    assert np.array_equal(a, b) and a.size == 3 << 20
    assert zlib.crc32(a.tobytes()) == zlib.crc32(corpus.mixed_entropy(3 << 20).tobytes())
    t = corpus.text_like(1 << 20)
    assert np.array_equal(t, corpus.text_like(1 << 20))
    assert 0.35 < len(zlib.compress(t.tobytes(), 6)) / t.size < 0.5
    assert np.array_equal(corpus.tail("mixed", 5 << 20, 1, 32768), corpus.mixed_entropy(5 << 20, corpus.SEED_MIXED + 0x10000)[-32768:])


def test_edge_case_set_has_the_survey_cases():
    from zlibk_b200 import corpus

    names = dict(corpus.edge_cases())
    for want in ("empty", "one", "run258", "run259", "n32767", "n32768", "n32769", "n65535", "n65536", "zeros_1MiB", "stale_window"):
        assert want in names
    assert names["stale_window"] == b"abcabc\0\0\0"


def test_plan_and_assemble():
    from zlibk_b200 import sharding

    spans = sharding.plan_shards(10_000_000, 4)
    assert spans[0][0] == 0 and spans[-1][1] == 10_000_000
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert all(s[0] % 32768 == 0 for s in spans)
    data = np.random.default_rng(1).integers(0, 256, 10_000_000, dtype=np.uint8).tobytes()
    metas = [(1000 + r, zlib.adler32(data[a:b]), 0, b - a) for r, (a, b) in enumerate(spans)]
    offsets, total, adler = sharding.assemble(metas)
    assert offsets == [2, 1002, 2003, 3005] and total == 2 + sum(m[0] for m in metas) + 6
    assert adler == zlib.adler32(data)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from zlibk_b200 import corpus, sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 6 << 20
    data = corpus.mixed_entropy(n).tobytes()
    a, b = sharding.plan_shards(n, world)[rank]
    shard = data[a:b]
    # stand-in for the per-rank GPU call: any body length, the true checksum of the shard
    meta = torch.tensor([len(zlib.compress(shard, 1)), zlib.adler32(shard), 0, len(shard)], dtype=torch.int64)
    g = sharding.exchange(meta, world)  # the one collective
    offsets, total, adler = sharding.assemble([tuple(r) for r in g.tolist()])
    q.put((rank, offsets, total, adler, zlib.adler32(data)))
    dist.destroy_process_group()


def test_two_rank_exchange_gloo():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, off0, tot0, ad0, want), (_, off1, tot1, ad1, _) = results
    assert off0 == off1 and tot0 == tot1 and ad0 == ad1 == want
    assert off0[0] == 2 and off0[1] > 2
