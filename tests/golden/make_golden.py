"""Generate tests/golden/vectors.json -- the committed known-answer fixtures of the hot path.

Two kinds of vectors, labelled by `source`:
  * "reference-test": the input and the expected answer are both stated in the reference's own tests
    (file:line in `cite`); the oracle is CHECKED against them here before anything is written.
  * "oracle": the expected answer is the output of the oracle (the C restatement of the reference algorithm),
    recorded so that the GPU path and later oracle revisions are compared against frozen bytes.  The Kotlin
    reference itself can not be run in this image (no Kotlin/Native, no JDK, no network), so these are NOT
    outputs of the reference binary -- DESIGN.md section 2 says what pins the oracle.

usage (from the repo root, CPU only):  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import random
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.bitio import BitWriter, fixed_litlen_lengths, zlib_header  # noqa: E402
from tests.oracle_binding import Oracle  # noqa: E402

oracle = Oracle(os.path.join(ROOT, "oracle", "libzk_oracle.so"))
T = "src/commonTest/kotlin/ai/solace/zlib/"

vectors = {"adler32": [], "inflate": [], "reference_deflate": []}

# ---- Adler-32 known answers stated in the reference's tests -------------------------------------------------
ADLER = [
    (b"", 0x00000001, T + "bitwise/test/Adler32UtilsTest.kt:113-180"),
    (b"a", 0x00620062, T + "bitwise/test/Adler32UtilsTest.kt:113-180"),
    (b"abc", 0x024D0127, T + "bitwise/test/Adler32UtilsTest.kt:113-180"),
    (b"message digest", 0x29750586, T + "bitwise/test/Adler32UtilsTest.kt:113-180"),
    (b"abcdefghijklmnopqrstuvwxyz", 0x90860B20, T + "bitwise/test/Adler32UtilsTest.kt:113-180"),
    (bytes(i & 0xFF for i in range(1000)), 0x1D03E73C, T + "bitwise/test/Adler32UtilsTest.kt:49-58"),
    (b"Wikipedia", 0x11E60398, T + "test/Adler32Test.kt:48-57"),
    (b"A", 66 * 65536 + 66, T + "test/Adler32IssueTest.kt:44-51"),
    (b"AB", 198 * 65536 + 132, T + "test/Adler32IssueTest.kt:44-51"),
]
for data, want, cite in ADLER:
    assert oracle.adler32(data) == want == zlib.adler32(data), cite
    vectors["adler32"].append({"source": "reference-test", "cite": cite, "data": data.hex(), "adler": want})
for n in (5551, 5552, 5553, 11104, 11000):  # Adler32Test.kt:37-44,174-236 (sizes around NMAX; values from the definition)
    data = bytes((i * 7 + 3) & 0xFF for i in range(n))
    a = b = 0
    a = 1
    for x in data:
        a = (a + x) % 65521
        b = (b + a) % 65521
    want = (b << 16) | a
    assert oracle.adler32(data) == want
    vectors["adler32"].append({"source": "reference-test", "cite": T + "test/Adler32Test.kt:37-44,174-236 (chunk-free definition)",
                               "gen": "bytes((i*7+3)&0xFF for i in range(%d))" % n, "n": n, "adler": want})

# ---- inflate: the reference's own test streams with their stated return codes ------------------------------
lit_codes, lit_bits = oracle.build_encoder(fixed_litlen_lengths())
dist_codes, dist_bits = oracle.build_encoder([5] * 32)


def add_inflate(name, stream, cite, want_rc=None, want_out=None):
    rc, out, second, consumed = oracle.inflate(stream, cap=1 << 20)
    if want_rc is not None:
        assert rc == want_rc, (name, rc)
    if want_out is not None:
        assert out == want_out, name
    vectors["inflate"].append({"source": "reference-test" if want_rc is not None else "oracle", "cite": cite, "name": name,
                               "stream": bytes(stream).hex(), "rc": rc, "out": out.hex(), "second": second, "consumed": consumed})


_, hello, _ = oracle.deflate(b"hello world")
add_inflate("hello world round trip", hello, T + "inflate/test/InflateBasicTests.kt:13-29", 1, b"hello world")
_, empty, _ = oracle.deflate(b"")
assert empty == bytes([0x78, 0x9C, 0x03, 0x00, 0x00, 0x00, 0x00, 0x01])
add_inflate("empty stream", empty, T + "deflate/DeflateParameterValidationTest.kt:32", 1, b"")
add_inflate("truncated after header", empty[:2], T + "inflate/test/InflateBasicTests.kt:31-48", -5, b"")
w = BitWriter().raw(zlib_header())
w.bits(1, 1).bits(1, 2).bits(lit_codes[257], lit_bits[257]).bits(dist_codes[0], dist_bits[0])
add_inflate("distance too far back", w.done(), T + "inflate/test/InflateErrorCasesTest.kt:17-60", -3, b"")
w = BitWriter().raw(zlib_header()).bits(1, 1).bits(0b11, 2)
add_inflate("reserved block type", w.done(), T + "inflate/test/InvalidBlockTypeTest.kt:11-29", -3)
add_inflate("FDICT set", zlib_header(flevel=2, fdict=1) + bytes([0x12, 0x34, 0x56, 0x78]), T + "inflate/test/InflateFdictTest.kt:11-38", 2, b"")
add_inflate("bad FCHECK", bytes([0x78, 0x9D, 3, 0, 0, 0, 0, 1]), "inflate/InflateStream.kt:28-35", -3)
add_inflate("CM != 8", bytes([0x79, 0x9C]), "inflate/InflateStream.kt:28-35", -3)
w = BitWriter().raw(zlib_header()).bits(1, 1).bits(0, 2).align()
w.raw(bytes([5, 0, 0xFF, 0xFF])).raw(b"hello")
add_inflate("stored LEN/NLEN mismatch", w.done(), "inflate/InflateStream.kt:47-74", -3)
z = bytearray(zlib.compress(b"some text some text some text", 6))
z[-1] ^= 1
add_inflate("Adler-32 mismatch", bytes(z), "inflate/InflateStream.kt:419-424", -3, b"some text some text some text")
# libz-made streams (the code pigz links): every block type, every level
rng = random.Random(11)
text = b" ".join(bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(1, 9))) for _ in range(600))
for level in (0, 1, 6, 9):
    add_inflate("libz level %d text" % level, zlib.compress(text, level), "libz 1.3 cross-check (SURVEY.md 8c)", 1, text)
add_inflate("libz fixed block", zlib.compress(b"abcabcabc", 9), "libz 1.3 cross-check (SURVEY.md 8c)", 1, b"abcabcabc")
for cut in (3, 7, 20):
    s = zlib.compress(text, 6)
    add_inflate("libz text cut %d bytes short" % cut, s[:-cut], "inflate/InflateStream.kt:414-428 (truncation)")

# ---- the reference's compressor (oracle restatement): frozen streams and what the reference inflate makes of them
CASES = [
    ("empty", b""), ("one byte", b"a"), ("hello world", b"hello world"), ("abcabc000", b"abcabc\0\0\0"),
    ("a x 1000", b"a" * 1000), ("text 4 KiB", text[:4096]), ("random 3000", random.Random(5).randbytes(3000)),
    ("zeros 70000", bytes(70000)),
]
for name, data in CASES:
    for level in (0, 1, 6, 9):
        rc, stream, used = oracle.deflate(data, level)
        irc, out, second, consumed = oracle.inflate(stream, cap=len(data) + 4096)
        vectors["reference_deflate"].append({
            "source": "oracle", "cite": "deflate/DeflateStream.kt:302-1045 restated in oracle/ref_deflate.c", "name": name, "level": level,
            "data": data.hex() if len(data) <= 4096 else None, "gen": None if len(data) <= 4096 else "bytes(70000)",
            "rc": rc, "bytes_in": used, "stream": stream.hex() if len(stream) <= 16384 else None,
            "stream_sha256": hashlib.sha256(stream).hexdigest(), "stream_len": len(stream),
            "inflate_rc": irc, "inflate_matches_input": out == data, "inflate_out_len": len(out), "inflate_consumed": consumed})

path = os.path.join(ROOT, "tests", "golden", "vectors.json")
with open(path, "w") as f:
    json.dump(vectors, f, indent=0, sort_keys=True)
print("wrote", path, {k: len(v) for k, v in vectors.items()}, os.path.getsize(path), "bytes")
