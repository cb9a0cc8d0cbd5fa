"""CRC-32 and the gzip / raw-deflate containers.  The reference has neither; libz 1.3 and Python's gzip
module (the formats pigz writes by default) are the oracle both ways."""
import gzip
import random
import zlib

import pytest

from tests.oracle_binding import Z_BUF_ERROR, Z_DATA_ERROR, Z_OK, Z_STREAM_END

pytestmark = pytest.mark.gpu


def _text(seed, n):
    rng = random.Random(seed)
    vocab = [bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(2, 9))) for _ in range(400)]
    out = bytearray()
    while len(out) < n:
        out += rng.choice(vocab) + b" "
    return bytes(out[:n])


@pytest.mark.parametrize("n", [0, 1, 2, 15, 16, 17, 255, 4095, 4096, 4097, 65535, 65536, 1_000_003, 9_000_001, 40_000_000])
def test_crc32_sizes_alignments_and_seeds(engine, n):
    data = random.Random(n).randbytes(n + 5)
    for off in (0, 1, 5):
        piece = data[off:off + n]
        assert engine.crc32(piece) == zlib.crc32(piece), (n, off)
    piece = data[:n]
    if n:
        cut = n // 3
        assert engine.crc32(piece[cut:], engine.crc32(piece[:cut])) == zlib.crc32(piece)  # incremental
        assert engine.crc32_combine(zlib.crc32(piece[:cut]), zlib.crc32(piece[cut:]), n - cut) == zlib.crc32(piece)
    assert engine.crc32(b"123456789") == 0xCBF43926  # the check value of CRC-32/ISO-HDLC


@pytest.mark.parametrize("level", [1, 6, 9])
def test_gzip_round_trips_both_ways(engine, level):
    for data in (b"", b"a", _text(1, 1000), _text(2, 300_000) + random.Random(3).randbytes(100_000) + bytes(200_000)):
        rc, stream, crc = engine.deflate_container(data, level, "gzip")
        stream = bytes(stream)
        assert rc == Z_OK and stream[:3] == b"\x1f\x8b\x08" and crc == zlib.crc32(data)
        assert gzip.decompress(stream) == data
        rc, out, used = engine.inflate_container(stream, len(data) + 16, "gzip")
        assert rc == Z_STREAM_END and bytes(out) == data and used == len(stream)
        # gzip-module output (FNAME absent, mtime set), and a member with FNAME + FCOMMENT + FEXTRA + FHCRC by hand
        theirs = gzip.compress(data, level, mtime=1234567)
        rc, out, used = engine.inflate_container(theirs, len(data) + 16, "gzip")
        assert rc == Z_STREAM_END and bytes(out) == data and used == len(theirs)
        body = zlib.compressobj(level, zlib.DEFLATED, -15)
        raw = body.compress(data) + body.flush()
        fancy = (b"\x1f\x8b\x08\x1e" + bytes(4) + b"\x00\x03" + b"\x04\x00abcd" + b"name.txt\x00" + b"a comment\x00" + b"\x12\x34" +
                 raw + zlib.crc32(data).to_bytes(4, "little") + (len(data) & 0xFFFFFFFF).to_bytes(4, "little"))
        rc, out, used = engine.inflate_container(fancy, len(data) + 16, "gzip")
        assert rc == Z_STREAM_END and bytes(out) == data and used == len(fancy)


def test_gzip_members_and_errors(engine):
    a, b = _text(4, 50_000), _text(5, 70_000)
    two = gzip.compress(a, 6, mtime=0) + bytes(engine.deflate_container(b, 6, "gzip")[1])
    rc, out, used = engine.inflate_container(two, len(a) + len(b) + 16, "gzip")
    assert rc == Z_STREAM_END and bytes(out) == a + b and used == len(two)  # RFC 1952 2.2: a series of members
    rc, out, used = engine.inflate_container(two + b"trailing garbage", len(a) + len(b) + 16, "gzip")
    assert rc == Z_STREAM_END and bytes(out) == a + b and used == len(two)
    one = bytearray(gzip.compress(a, 6, mtime=0))
    one[-5] ^= 1  # CRC-32
    assert engine.inflate_container(bytes(one), len(a) + 16, "gzip")[0] == Z_DATA_ERROR
    one[-5] ^= 1
    one[-1] ^= 1  # ISIZE
    assert engine.inflate_container(bytes(one), len(a) + 16, "gzip")[0] == Z_DATA_ERROR
    one[-1] ^= 1
    assert engine.inflate_container(bytes(one[:-3]), len(a) + 16, "gzip")[0] == Z_BUF_ERROR
    assert engine.inflate_container(b"\x1f\x8c\x08" + bytes(20), 100, "gzip")[0] == Z_DATA_ERROR


@pytest.mark.parametrize("level", [1, 6])
def test_raw_deflate_both_ways(engine, level):
    data = _text(6, 500_000)
    rc, stream, _ = engine.deflate_container(data, level, "raw")
    assert rc == Z_OK and zlib.decompress(bytes(stream), -15) == data
    rc, out, used = engine.inflate_container(stream, len(data), "raw")
    assert rc == Z_STREAM_END and bytes(out) == data and used == len(stream)
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    theirs = c.compress(data) + c.flush()
    rc, out, used = engine.inflate_container(theirs, len(data), "raw")
    assert rc == Z_STREAM_END and bytes(out) == data and used == len(theirs)


def test_gzip_large_through_the_lane_decoder(engine):
    import numpy as np

    from zlibk_b200 import corpus

    data = corpus.mixed_entropy(48 << 20)
    rc, stream, crc = engine.deflate_container(data, 6, "gzip")
    assert rc == Z_OK and crc == zlib.crc32(data)
    assert gzip.decompress(bytes(stream)) == data.tobytes()
    rc, out, used = engine.inflate_container(stream, data.size, "gzip")
    assert rc == Z_STREAM_END and used == len(stream) and np.array_equal(np.asarray(out), data)


def _zstream_deflate(zk, data_pieces, windowBits, level=6, copy_after=None):
    z = zk.ZStream()
    assert z.deflateInit(level, 8, windowBits) == zk.Z_OK
    out, twin = bytearray(), None
    for i, piece in enumerate(data_pieces):
        z.next_in, z.next_in_index, z.avail_in = piece, 0, len(piece)
        last = i == len(data_pieces) - 1
        while True:
            z.next_out, z.next_out_index, z.avail_out = bytearray(1 << 16), 0, 1 << 16
            rc = z.deflate(zk.Z_FINISH if last else zk.Z_SYNC_FLUSH)
            out += z.next_out[: z.next_out_index]
            if rc == zk.Z_STREAM_END or (not last and z.avail_in == 0 and z.avail_out > 0):
                break
            assert rc in (zk.Z_OK, zk.Z_BUF_ERROR)
        if copy_after == i:
            twin = zk.ZStream()
            assert z.deflateCopy(twin) == zk.Z_OK
    z.deflateEnd()
    return bytes(out), twin


def test_zstream_window_bits_select_the_container(engine):
    """docs/API.md:239-281 with zlib.h's windowBits convention: negative = raw deflate, +16 = gzip."""
    import zlibk_b200 as zk

    data = _text(7, 400_000)
    pieces = [data[:1000], data[1000:250_000], data[250_000:]]
    gz, _ = _zstream_deflate(zk, pieces, 31)
    assert gz[:2] == b"\x1f\x8b" and gzip.decompress(gz) == data
    raw, _ = _zstream_deflate(zk, pieces, -15)
    assert zlib.decompress(raw, -15) == data
    zl, _ = _zstream_deflate(zk, pieces, 15)
    assert zlib.decompress(zl) == data
    for stream, wbits in ((gzip.compress(data, 6, mtime=0), 31), (raw, -15), (zlib.compress(data, 6), 15)):
        y = zk.ZStream()
        assert y.inflateInit(wbits) == zk.Z_OK
        y.next_in, y.next_in_index, y.avail_in = stream, 0, len(stream)
        y.next_out, y.next_out_index, y.avail_out = bytearray(len(data)), 0, len(data)
        assert y.inflate(zk.Z_FINISH) == zk.Z_STREAM_END
        assert bytes(y.next_out[: y.next_out_index]) == data
        y.inflateEnd()
    assert zk.ZStream().deflateInit(6, 8, 7) == zk.Z_STREAM_ERROR
    assert zk.ZStream().inflateInit(40) == zk.Z_STREAM_ERROR


def test_zstream_deflate_copy(engine):
    """docs/API.md:253: after deflateCopy both streams continue independently from the same state."""
    import zlibk_b200 as zk

    head, tail_a, tail_b = _text(8, 200_000), _text(9, 100_000), _text(10, 50_000)
    a, twin = _zstream_deflate(zk, [head, tail_a], 15, copy_after=0)
    assert zlib.decompress(a) == head + tail_a
    twin.next_in, twin.next_in_index, twin.avail_in = tail_b, 0, len(tail_b)
    twin.next_out, twin.next_out_index, twin.avail_out = bytearray(1 << 20), 0, 1 << 20
    assert twin.deflate(zk.Z_FINISH) == zk.Z_STREAM_END
    stream_b = bytes(twin.next_out[: twin.next_out_index])
    twin.deflateEnd()
    # the twin's output continues the original's output at the copy point: header + first piece come from the original
    z = zk.ZStream()
    assert z.deflateInit(6) == zk.Z_OK
    z.next_in, z.next_in_index, z.avail_in = head, 0, len(head)
    z.next_out, z.next_out_index, z.avail_out = bytearray(1 << 20), 0, 1 << 20
    assert z.deflate(zk.Z_SYNC_FLUSH) == zk.Z_OK
    prefix = bytes(z.next_out[: z.next_out_index])
    z.deflateEnd()
    assert zlib.decompress(prefix + stream_b) == head + tail_b


@pytest.mark.gpu
def test_level_zero_in_raw_and_gzip_containers_is_stored(engine):
    """level 0 means stored blocks in every container (DS:307-308), not a silent level 1."""
    data = _text(12, 300_000) + bytes(range(256)) * 300
    for level in (0, -1):
        rc, raw, _ = engine.deflate_container(data, level, container="raw")
        assert rc == 0 and zlib.decompress(bytes(raw), -15) == data
        assert len(raw) >= len(data)  # stored: no compression at all
        assert bytes(raw)[0] & 6 == 0  # BTYPE = 00
        rc, gz, crc = engine.deflate_container(data, level, container="gzip")
        assert rc == 0 and gzip.decompress(bytes(gz)) == data and crc == zlib.crc32(data)
        assert len(gz) >= len(data)
        rc, back, used = engine.inflate_container(gz, len(data), container="gzip")
        assert rc == 1 and bytes(back) == data and used == len(gz)
