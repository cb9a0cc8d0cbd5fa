"""Preset dictionaries (zlib.h deflateSetDictionary / inflateSetDictionary).  The reference only
detects FDICT and answers Z_NEED_DICT (inflate/InflateStream.kt:36-43) -- that behaviour is kept when no dictionary is
supplied -- so libz 1.3 is the oracle for the rest: its FDICT streams must decode on the GPU, and the GPU's must decode in
libz."""
import random
import zlib

import pytest

from tests.oracle_binding import Z_DATA_ERROR, Z_NEED_DICT, Z_OK, Z_STREAM_END

pytestmark = pytest.mark.gpu


def _words(seed, n):
    rng = random.Random(seed)
    vocab = [bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(2, 9))) for _ in range(300)]
    out = bytearray()
    while len(out) < n:
        out += rng.choice(vocab) + b" "
    return bytes(out[:n])


def _libz_dict_stream(data, zdict, level=6):
    c = zlib.compressobj(level, zdict=zdict)
    return c.compress(data) + c.flush()


@pytest.mark.parametrize("n,dn", [(0, 10), (1, 1), (1000, 300), (50_000, 32768), (200_000, 40_000), (70_000, 5)])
def test_libz_fdict_streams_decode_on_the_gpu(engine, oracle, n, dn):
    zdict = _words(1, dn)
    data = zdict[-min(dn, 200):] * 3 + _words(1, n)  # starts with references into the dictionary
    stream = _libz_dict_stream(data, zdict)
    assert stream[1] & 0x20
    # without a dictionary: exactly what the reference does (IS:36-43)
    rc, out, used = engine.inflate(stream, len(data) + 64)
    orc, oout, _, oused = oracle.inflate(stream, cap=len(data) + 64)
    assert (rc, bytes(out), used) == (orc, oout, oused) == (Z_NEED_DICT, b"", 6)
    rc, out, used = engine.inflate_dict(stream, zdict, len(data) + 64)
    assert rc == Z_STREAM_END and bytes(out) == data and used == len(stream)
    rc, out, _ = engine.inflate_dict(stream, zdict[:-1] + b"?", len(data) + 64)
    assert rc == Z_DATA_ERROR  # DICTID mismatch


@pytest.mark.parametrize("level", [1, 6, 9])
def test_gpu_fdict_streams_decode_in_libz(engine, level):
    zdict = _words(2, 50_000)  # only the last 32 KiB can be referenced
    data = _words(2, 300_000)
    rc, stream, adler = engine.deflate_dict(zdict, data, level)
    stream = bytes(stream)
    assert rc == Z_OK and stream[0] == 0x78 and stream[1] & 0x20 and ((stream[0] << 8) | stream[1]) % 31 == 0
    assert int.from_bytes(stream[2:6], "big") == zlib.adler32(zdict)
    d = zlib.decompressobj(zdict=zdict)
    assert d.decompress(stream) + d.flush() == data and d.eof
    assert adler == zlib.adler32(data)
    rc, plain, _ = engine.deflate(data, level)
    assert len(stream) < len(plain)  # the dictionary pays (same vocabulary)
    rc, out, used = engine.inflate_dict(stream, zdict, len(data))
    assert rc == Z_STREAM_END and bytes(out) == data and used == len(stream)


def test_dictionary_through_the_lane_decoder(engine):
    """16 MiB: more than 256 blocks, so the lane-per-block decoder runs, with block 0 reaching into the dictionary."""
    import numpy as np

    zdict = _words(3, 32768)
    data = (zdict[-3000:] + _words(4, 1 << 20)) * 16
    rc, stream, _ = engine.deflate_dict(zdict, data, 6)
    assert rc == Z_OK
    d = zlib.decompressobj(zdict=zdict)
    assert d.decompress(bytes(stream)) == data
    rc, out, used = engine.inflate_dict(stream, zdict, len(data))
    assert rc == Z_STREAM_END and np.array_equal(np.frombuffer(data, dtype=np.uint8), np.asarray(out)) and used == len(stream)
    # a reference that reaches before the dictionary is a data error, as a reference before the output is (IS:287-291)
    rc, out, used = engine.inflate_dict(stream, zdict[-100:], len(data))
    assert rc == Z_DATA_ERROR


def test_zstream_dictionary_calls(engine):
    """docs/API.md:239-281: deflateSetDictionary before the first deflate; inflate answers Z_NEED_DICT, then
    inflateSetDictionary and inflate again."""
    import zlibk_b200 as zk

    zdict, data = _words(5, 4000), _words(5, 100_000)
    z = zk.ZStream()
    assert z.deflateInit(6) == zk.Z_OK
    assert z.deflateSetDictionary(zdict) == zk.Z_OK
    z.next_in, z.next_in_index, z.avail_in = data, 0, len(data)
    z.next_out, z.next_out_index, z.avail_out = bytearray(200_000), 0, 200_000
    assert z.deflate(zk.Z_FINISH) == zk.Z_STREAM_END
    stream = bytes(z.next_out[: z.next_out_index])
    z.deflateEnd()
    d = zlib.decompressobj(zdict=zdict)
    assert d.decompress(stream) == data
    y = zk.ZStream()
    assert y.inflateInit() == zk.Z_OK
    y.next_in, y.next_in_index, y.avail_in = stream, 0, len(stream)
    y.next_out, y.next_out_index, y.avail_out = bytearray(len(data)), 0, len(data)
    assert y.inflate(zk.Z_NO_FLUSH) == zk.Z_NEED_DICT
    assert y.inflateSetDictionary(zdict) == zk.Z_OK
    assert y.inflate(zk.Z_FINISH) == zk.Z_STREAM_END
    assert bytes(y.next_out[: y.next_out_index]) == data
    y.inflateEnd()
