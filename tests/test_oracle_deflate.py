"""Pin the oracle's restatement of the reference compressor with what the reference's
tests hold (deflate/test/DeflateParameterValidationTest.kt:12-44,
inflate/test/InflateBasicTests.kt:13-29) and document its known defects.
Compressed BYTES of the reference are not pinned by any reference test: "parity
unpinned" for deflate output beyond these."""
import random
import zlib

import pytest

from tests.oracle_binding import Z_DATA_ERROR, Z_OK, Z_STREAM_END, Z_STREAM_ERROR


def test_level_above_max_writes_nothing(oracle):
    assert oracle.deflate(b"", 10) == (Z_STREAM_ERROR, b"", 0)
    assert oracle.deflate(b"abc", 11) == (Z_STREAM_ERROR, b"", 0)


def test_level_nine_empty_input(oracle):
    rc, s, used = oracle.deflate(b"", 9)
    assert (rc, used) == (Z_OK, 0) and len(s) >= 6
    assert s == bytes.fromhex("78da030000000001")
    assert zlib.decompress(s) == b""


def test_empty_input_level_six_shape(oracle):
    assert oracle.deflate(b"", 6)[1] == bytes.fromhex("789c030000000001")


def test_negative_level_is_stored(oracle):
    rc, s, used = oracle.deflate(b"", -1)
    assert (rc, used) == (Z_OK, 0) and len(s) >= 6
    assert s[:2] == bytes([0x78, 0x01]) and s[2] == 0x01  # BFINAL=1, BTYPE=00
    assert zlib.decompress(s) == b""


@pytest.mark.parametrize("level,flg", [(0, 0x01), (1, 0x01), (2, 0x5E), (5, 0x5E), (6, 0x9C), (8, 0x9C), (9, 0xDA)])
def test_header_bytes(oracle, level, flg):
    assert oracle.deflate(b"x", level)[1][:2] == bytes([0x78, flg])  # DeflateStream.kt:27-50


@pytest.mark.parametrize("level", [0, 1, 2, 6, 9])
def test_small_inputs_round_trip(oracle, level):
    for data in (b"hello world", b"a", b"ab", b"abc", b"a" * 1000, bytes(range(256)) * 4, b"the quick brown fox " * 40):
        rc, s, used = oracle.deflate(data, level)
        assert (rc, used) == (Z_OK, len(data))
        assert oracle.inflate(s)[:2] == (Z_STREAM_END, data)
        assert zlib.decompress(s) == data


def test_a_times_1000_size(oracle):
    # 'a' x 1000 -> 61 bytes at level 6 (libz: 17)
    assert len(oracle.deflate(b"a" * 1000, 6)[1]) == 61


def test_length_258_uses_symbol_284(oracle):
    # DeflateStream.kt:193-203: 258 maps to 284 + 5 extra bits, still valid for every inflater
    data = b"\x00" * 4000
    s = oracle.deflate(data, 6)[1]
    assert zlib.decompress(s) == data


def test_stale_window_defect_is_reproduced(oracle):
    # SURVEY.md section 0 item 5 / BASELINE.md: "abcabc\0\0\0" -> 12 bytes that decode to "abcabcabc"
    rc, s, _ = oracle.deflate(b"abcabc\0\0\0", 6)
    assert rc == Z_OK and len(s) == 12
    rc, out, _, _ = oracle.inflate(s)
    assert rc == Z_DATA_ERROR and out == b"abcabcabc"  # Adler trailer mismatch after wrong bytes


def test_stored_path_defect_is_reproduced(oracle):
    # DeflateStream.kt:350 reads every Okio segment to offset 0: inputs longer than one
    # 8 KiB segment yield a self-consistent stream of the WRONG bytes.
    data = random.Random(11).randbytes(20000)
    s = oracle.deflate(data, 0)[1]
    rc, out, _, _ = oracle.inflate(s)
    assert rc == Z_STREAM_END and len(out) == len(data) and out != data
    small = data[:8000]
    assert oracle.inflate(oracle.deflate(small, 0)[1])[1] == small


def test_random_data_falls_back_to_stored_blocks(oracle):
    data = random.Random(12).randbytes(150_000)
    s = oracle.deflate(data, 6)[1]
    assert len(s) == len(data) + 2 + 4 + 5 * 3  # three stored blocks (+21 B / 150 KB)
    assert zlib.decompress(s) == data


def test_repetitive_lines_round_trip_and_ratio(oracle):
    lines = b"".join(b"2024-01-01 12:00:%02d INFO request id=%06d status=200 OK\n" % (i % 60, i) for i in range(4000))
    s = oracle.deflate(lines, 6)[1]
    assert oracle.inflate(s, cap=len(lines) + 8)[:2] == (Z_STREAM_END, lines)
    assert len(s) < 0.2 * len(lines)
