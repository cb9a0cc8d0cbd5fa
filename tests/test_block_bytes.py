"""The "block_bytes" option against incompressible input: every stored block costs a header and a sync marker, so the
bound must hold for the smallest block span (n + n/2048 + 64 did not), in every entry point that sizes its
output from zk_compress_bound."""
import zlib

import numpy as np
import pytest


def test_bound_covers_smallest_blocks():
    import zlibk_b200

    lib = zlibk_b200.library()
    for n in (0, 1, 4095, 4096, 4097, 1 << 20, (1 << 30) + 5):
        blocks = n // 4096 + 1
        assert lib.zk_compress_bound(n) >= n + 10 * blocks + 2 + 2 + 4


@pytest.mark.gpu
@pytest.mark.parametrize("block_bytes", [4096, 8192, 16384, 32768])
def test_random_data_fits_the_bound(engine, block_bytes):
    rng = np.random.default_rng(block_bytes)
    data = rng.integers(0, 256, size=(3 << 20) + 777, dtype=np.uint8)
    engine.set_option("block_bytes", block_bytes)
    try:
        for level in (1, 6):
            rc, stream, adler = engine.deflate(data, level)
            assert rc == 0, (block_bytes, level, rc)
            assert len(stream) <= engine.compress_bound(len(data))
            assert zlib.decompress(bytes(stream)) == data.tobytes()
            rc, body, _ = engine.deflate_shard(None, data, level)
            assert rc == 0
            rc, raw, _ = engine.deflate_container(data, level, container="raw")
            assert rc == 0 and zlib.decompress(bytes(raw), -15) == data.tobytes()
    finally:
        engine.set_option("block_bytes", 32768)


@pytest.mark.gpu
def test_host_slab_option_is_bounded(engine):
    with pytest.raises(ValueError):
        engine.set_option("host_slab_bytes", (1 << 30) + 65536)
    with pytest.raises(ValueError):
        engine.set_option("host_first_slab_bytes", 2 << 30)
    engine.set_option("host_slab_bytes", 0)
