"""The pipelined host-buffer inflate (zk_api.cu host_inflate_piped, opt-in with the "inflate_pipe" option): slabs cut at
sync markers, decoded as shard bodies on worker contexts, finished in order, copied down while the next slabs decode.  It
must be indistinguishable from the one-pass path: same bytes, same return code, same consumed count -- and anything it cannot handle must reach the
one-pass path, which owns the reference's error behaviour (IS:425-437)."""
import zlib

import numpy as np
import pytest

from zlibk_b200 import Z_BUF_ERROR, Z_DATA_ERROR, Z_OK, Z_STREAM_END


@pytest.fixture(scope="module")
def big_stream(engine):
    from zlibk_b200 import corpus

    data = corpus.mixed_entropy(320 << 20)
    rc, stream, _ = engine.deflate(data, 6)
    assert rc == Z_OK and len(stream) > (100 << 20)
    return data, np.array(stream, copy=True)


@pytest.mark.gpu
def test_piped_equals_one_pass(engine, big_stream):
    data, stream = big_stream
    n = len(data)
    engine.set_option("inflate_pipe", 0)
    rc0, out0, used0 = engine.inflate(stream, n)
    engine.set_option("inflate_pipe", 1)
    try:
        for slab in (0, 16 << 20, 64 << 20):
            engine.set_option("inflate_slab_bytes", slab)
            before = engine.counter("inflate_pipe_slabs")
            rc, out, used = engine.inflate(stream, n)
            assert engine.counter("inflate_pipe_slabs") > before, "the pipeline did not run"
            assert (rc, used) == (rc0, used0) == (Z_STREAM_END, len(stream))
            assert np.array_equal(np.asarray(out), np.asarray(out0)) and np.array_equal(np.asarray(out), data)
    finally:
        engine.set_option("inflate_slab_bytes", 0)
        engine.set_option("inflate_pipe", 0)


@pytest.mark.gpu
def test_piped_errors_fall_back_to_the_reference_behaviour(engine, big_stream, oracle):
    data, stream = big_stream
    n = len(data)
    # a flipped bit in the trailer: every byte comes out, the checksum does not match (IS:419-424)
    bad = stream.copy()
    bad[-1] ^= 1
    engine.set_option("inflate_pipe", 1)
    rc, out, used = engine.inflate(bad, n)
    assert rc == Z_DATA_ERROR and len(out) == n and np.array_equal(np.asarray(out), data)
    # damage inside the stream (a Huffman block near the start, a block in the middle, the last marker): whatever the
    # one-pass path answers is the answer -- directly, or through the fallback
    fb = engine.counter("inflate_pipe_fallbacks")
    for at in (5000, len(stream) // 2, len(stream) - 9):
        bad = stream.copy()
        bad[at: at + 8] ^= 0x5A
        engine.set_option("inflate_pipe", 0)
        want = engine.inflate(bad, n)
        engine.set_option("inflate_pipe", 1)
        got = engine.inflate(bad, n)
        assert got[0] == want[0] and got[2] == want[2], (at, got[0], want[0], got[2], want[2])
        assert np.array_equal(np.asarray(got[1]), np.asarray(want[1])), at
    assert engine.counter("inflate_pipe_fallbacks") > fb  # the broken Huffman block cannot pass as a shard body
    # output buffer too small, truncated input: not this path's business
    engine.set_option("inflate_pipe", 1)
    rc, out, _ = engine.inflate(stream, n - 4096)
    assert rc == Z_BUF_ERROR
    rc, out, _ = engine.inflate(stream[: len(stream) - 100], n)
    assert rc == Z_BUF_ERROR
    # a libz stream of the same data has no markers: one pass, same answer as ever
    z = zlib.compress(data[: 80 << 20].tobytes(), 1)
    rc, out, used = engine.inflate(z, 80 << 20)
    assert rc == Z_STREAM_END and used == len(z) and np.array_equal(np.asarray(out), data[: 80 << 20])
    engine.set_option("inflate_pipe", 0)
