"""Pin the oracle's inflate against the reference's own inflate tests (exact return
codes) and cross-check it against libz 1.3 on libz-made streams.

Reference tests restated here:
  inflate/test/InflateBasicTests.kt:13-48, InflateErrorCasesTest.kt:17-60,
  InvalidBlockTypeTest.kt:11-29, InflateFdictTest.kt:11-38,
  inflate/CanonicalHuffmanFullTableInvariantTest.kt:8-57
"""
import random
import zlib

import pytest

from tests.bitio import BitWriter, fixed_litlen_lengths, zlib_header
from tests.oracle_binding import (ESCAPED_EXCEPTION, Z_BUF_ERROR, Z_DATA_ERROR, Z_NEED_DICT,
                                  Z_STREAM_END)


def test_hello_world_round_trip(oracle):
    rc, stream, used = oracle.deflate(b"hello world")  # default level 6
    assert (rc, used) == (0, 11)
    rc, out, second, consumed = oracle.inflate(stream)
    assert (rc, out, second) == (Z_STREAM_END, b"hello world", 11)  # Z_STREAM_END, not Z_OK
    assert consumed == len(stream)
    assert zlib.decompress(stream) == b"hello world"


def test_truncated_after_header_is_buf_error(oracle):
    _, full, _ = oracle.deflate(b"")
    rc, out, _, _ = oracle.inflate(full[:2])
    assert rc == Z_BUF_ERROR and out == b""


def test_distance_too_far_back_is_data_error(oracle):
    lit_codes, lit_bits = oracle.build_encoder(fixed_litlen_lengths())
    dist_codes, dist_bits = oracle.build_encoder([5] * 32)
    w = BitWriter().raw(zlib_header())
    w.bits(1, 1).bits(1, 2)
    w.bits(lit_codes[257], lit_bits[257])  # length 3 before any literal
    w.bits(dist_codes[0], dist_bits[0])  # distance 1
    rc, out, _, _ = oracle.inflate(w.done())
    assert rc == Z_DATA_ERROR and out == b""


def test_reserved_block_type_is_data_error(oracle):
    w = BitWriter().raw(zlib_header()).bits(1, 1).bits(0b11, 2)
    rc, _, _, _ = oracle.inflate(w.done())
    assert rc == Z_DATA_ERROR


def test_fdict_is_need_dict_with_no_output(oracle):
    stream = zlib_header(flevel=2, fdict=1) + bytes([0x12, 0x34, 0x56, 0x78])
    rc, out, second, consumed = oracle.inflate(stream)
    assert (rc, out, second) == (Z_NEED_DICT, b"", 0)
    assert consumed == 6  # DICTID is consumed (InflateStream.kt:41)


def test_bad_header_is_data_error(oracle):
    assert oracle.inflate(bytes([0x78, 0x9D, 3, 0, 0, 0, 0, 1]))[0] == Z_DATA_ERROR  # FCHECK
    assert oracle.inflate(bytes([0x79, 0x9C]))[0] == Z_DATA_ERROR  # CM != 8
    assert oracle.inflate(bytes([0x88, 0x1C]))[0] == Z_DATA_ERROR  # CINFO 8


def test_full_table_invariants(oracle):
    assert oracle.build_full([16])[0] == -1  # buildFull rejects length 16
    m, bits, vals = oracle.build_full([1, 1])
    assert m == 1 and bits == [1, 1] and vals == [0, 1]
    m, bits, vals = oracle.build_full(fixed_litlen_lengths())
    assert m == 9 and len(bits) == 512 and all(b in (7, 8, 9) for b in bits)


def test_stored_len_nlen_mismatch(oracle):
    w = BitWriter().raw(zlib_header()).bits(1, 1).bits(0, 2).align()
    w.raw(bytes([5, 0, 0xFF, 0xFF])).raw(b"hello")
    assert oracle.inflate(w.done())[0] == Z_DATA_ERROR


def test_adler_mismatch_is_data_error(oracle):
    z = bytearray(zlib.compress(b"some text some text some text", 6))
    z[-1] ^= 1
    rc, out, second, _ = oracle.inflate(bytes(z))
    assert rc == Z_DATA_ERROR and out == b"some text some text some text" and second == len(out)


def test_empty_code_length_table_escapes(oracle):
    # dynamic block whose 19 code-length-code lengths are all zero: the reference's
    # readCodeLengths -> decodeOne raises IllegalStateException("Empty Huffman table"),
    # which inflateZlib does not catch (InflateStream.kt:243, CanonicalHuffman.kt:127)
    w = BitWriter().raw(zlib_header()).bits(1, 1).bits(2, 2).bits(0, 5).bits(0, 5).bits(0, 4)
    for _ in range(4):
        w.bits(0, 3)
    w.raw  # noqa: B018
    rc = oracle.inflate(w.done() + bytes(8))[0]
    assert rc == ESCAPED_EXCEPTION


def _corpus(seed, n):
    rng = random.Random(seed)
    words = [bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(1, 10))) for _ in range(800)]
    out = bytearray()
    while len(out) < n:
        out += rng.choice(words) + b" "
    return bytes(out[:n])


@pytest.mark.parametrize("level", [0, 1, 6, 9])
@pytest.mark.parametrize("kind", ["text", "random", "zeros", "mixed"])
def test_libz_streams_decode_like_libz(oracle, level, kind):
    n = 200_000
    if kind == "text":
        data = _corpus(1, n)
    elif kind == "random":
        data = random.Random(2).randbytes(n)
    elif kind == "zeros":
        data = bytes(n)
    else:
        data = _corpus(3, 70_000) + random.Random(4).randbytes(60_000) + bytes(70_000)
    stream = zlib.compress(data, level)
    rc, out, second, consumed = oracle.inflate(stream, cap=n + 16)
    assert rc == Z_STREAM_END and out == data and second == n and consumed == len(stream)


def test_pigz_style_stream(oracle):
    """pigz -z framing rebuilt with libz: independent chunks primed with the previous
    32 KiB as dictionary, each ended by Z_SYNC_FLUSH (docs/CLI_PIGZ_TESTING.md; the pigz
    binary is not in the image)."""
    data = _corpus(5, 700_000)
    chunk = 128 * 1024
    body = bytearray()
    for off in range(0, len(data), chunk):
        c = zlib.compressobj(6, zlib.DEFLATED, -15, zdict=data[max(0, off - 32768):off]) if off else zlib.compressobj(6, zlib.DEFLATED, -15)
        body += c.compress(data[off:off + chunk])
        last = off + chunk >= len(data)
        body += c.flush(zlib.Z_FINISH if last else zlib.Z_SYNC_FLUSH)
    stream = zlib_header() + bytes(body) + zlib.adler32(data).to_bytes(4, "big")
    assert zlib.decompress(stream) == data
    rc, out, _, _ = oracle.inflate(stream, cap=len(data) + 16)
    assert rc == Z_STREAM_END and out == data


def test_trailing_bytes_are_left_unread(oracle):
    s = zlib.compress(b"first member") + zlib.compress(b"second")
    rc, out, _, consumed = oracle.inflate(s)
    assert rc == Z_STREAM_END and out == b"first member"
    assert consumed <= len(zlib.compress(b"first member")) + 4  # bit reader may prefetch a few bytes
