"""BASELINE.json configs[0]: the reference's CPU path on a 64 MiB text-like buffer at level 6, as a round trip.

The honest record (SURVEY.md section 0 item 5, 8d "expect and report reference failure"): the reference's compressor
(oracle restatement of DeflateStream.kt with its stale-window compare, DS:672-691) accepts the buffer and returns Z_OK, but
the stream it writes is rejected by the reference's own inflate at the first block -- Z_DATA_ERROR with no bytes out -- and
by libz ("invalid code lengths set").  libz round-trips the same buffer.  bench.py reports the same leg, with timings, as
cpu_baseline.config1.
"""
import zlib

import pytest

from tests.oracle_binding import Z_DATA_ERROR, Z_OK


@pytest.fixture(scope="module")
def text64():
    import zlibk_b200
    from zlibk_b200 import corpus

    return corpus.text_like(64 << 20).tobytes()


def test_config1_reference_round_trip_outcome(oracle, text64):
    rc, stream, used = oracle.deflate(text64, 6)
    assert (rc, used) == (Z_OK, len(text64))
    assert stream[:2] == bytes([0x78, 0x9C])
    assert 0.40 < len(stream) / len(text64) < 0.50  # the size of a stream nobody can read
    rc_i, out, second, _ = oracle.inflate(stream, cap=len(text64) + 64)
    assert rc_i == Z_DATA_ERROR and len(out) == 0
    with pytest.raises(zlib.error):
        zlib.decompress(stream)


def test_config1_libz_round_trip(text64):
    z = zlib.compress(text64, 6)
    assert zlib.decompress(z) == text64
    assert 0.39 < len(z) / len(text64) < 0.42
