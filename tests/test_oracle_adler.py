"""Pin the oracle's Adler-32 against every known-answer value in the reference's tests
(src/commonTest/.../Adler32UtilsTest.kt, Adler32EngineTest.kt,
test/Adler32Test.kt, test/Adler32IssueTest.kt, test/VerifyLargeBufferTest.kt) and libz."""
import random
import zlib

import pytest

KATS = [
    (b"", 0x00000001),  # Adler32UtilsTest.kt:113-180
    (b"a", 0x00620062),
    (b"abc", 0x024D0127),
    (b"message digest", 0x29750586),
    (b"abcdefghijklmnopqrstuvwxyz", 0x90860B20),
    (bytes(i & 0xFF for i in range(1000)), 0x1D03E73C),  # Adler32UtilsTest.kt:49-58
    (b"Wikipedia", 0x11E60398),  # Adler32Test.kt:48-57
    (b"A", 66 * 65536 + 66),  # Adler32IssueTest.kt:44-51
    (b"AB", 198 * 65536 + 132),
]


@pytest.mark.parametrize("data,want", KATS)
def test_known_answers(oracle, data, want):
    assert oracle.adler32(data) == want
    assert zlib.adler32(data) == want


def test_null_buffer_resets_to_one(oracle):
    # Adler32Test.kt:20-21: null buffer -> 1 even with seed 0
    assert oracle.adler32(None, adler=0) == 1
    assert oracle.adler32(None, adler=0x12345678) == 1


def test_incremental_equals_one_shot(oracle):
    # Adler32UtilsTest.kt:61-78,94-110
    data = bytes((i * 7 + 3) & 0xFF for i in range(20000))
    whole = oracle.adler32(data)
    for cut in (1, 100, 5551, 5552, 5553, 11104, 19999):
        part = oracle.adler32(data[:cut])
        assert oracle.adler32(data[cut:], adler=part) == whole
    assert whole == zlib.adler32(data)


@pytest.mark.parametrize("n", [5551, 5552, 5553, 2 * 5552, 11000, 5000])
def test_nmax_boundaries(oracle, n):
    # Adler32Test.kt:37-44,174-236: NMAX-1 / NMAX / NMAX+1 / 2*NMAX / 11000 / 5000 x 'a'
    for data in (b"a" * n, bytes([0xFF]) * n, bytes(random.Random(n).randrange(256) for _ in range(n))):
        a, b = 1, 0
        for x in data:  # the chunk-free definition
            a = (a + x) % 65521
            b = (b + a) % 65521
        assert oracle.adler32(data) == (b << 16) | a == zlib.adler32(data)


def test_large_buffer_matches_libz(oracle):
    # VerifyLargeBufferTest.kt:11-23 / Adler32Test.kt:238-257 (1 MB, log only upstream)
    data = bytes(random.Random(7).randrange(256) for _ in range(1_000_000))
    assert oracle.adler32(data) == zlib.adler32(data)


def test_version(oracle):
    assert oracle.version() == "1.0.2"  # common/Constants.kt:37-44
