"""BASELINE.json config 5 / SURVEY.md 8(f) row 2: the CLI with the reference's commands and messages, and the
pigz matrix of docs/CLI_PIGZ_TESTING.md:98-245 on the files scripts/test_cli_pigz_integration.sh:62-106 generates.
The pigz binary is not in the image: libz 1.3 (the code pigz links) plays it, including pigz's own framing
(independent chunks primed with the previous 32 KiB, joined by sync flushes)."""
import io
import os
import random
import zlib

import pytest

from tests.bitio import zlib_header

pytestmark = pytest.mark.gpu


def _script_files():
    rng = random.Random(0xC11)
    files = {"small.txt": b"Hello World! This is a small test file for ZLib compression.\n"}
    files["medium_repetitive.txt"] = b"".join(
        b"Line %d: This is repetitive content that should compress well with zlib format.\n" % i for i in range(1, 101))
    diverse = [b"ZLib.kotlin CLI Testing\n=======================\n\n",
               b"This file contains diverse text content to test compression efficiency.\n",
               b"It includes special characters: !@#$%^&*()_+-=[]{}|;:,.<>?\nAnd some numbers: 1234567890\n\n"]
    diverse += [b"Diverse line %d with random content %09d\n" % (i, rng.randrange(10 ** 9)) for i in range(1, 51)]
    files["medium_diverse.txt"] = b"".join(diverse)
    large = [b"Large Test File for ZLib.kotlin\n===============================\n\n"]
    for i in range(1, 1001):
        if i % 3 == 0:
            large.append(b"Repetitive content line %d: AAAAAABBBBBBCCCCCCDDDDDD\n" % i)
        elif i % 3 == 1:
            large.append(b"Mixed content line %d: %s\n" % (i, rng.randbytes(32).hex().encode()))
        else:
            large.append(b"Standard line %d: The quick brown fox jumps over the lazy dog.\n" % i)
    files["large_mixed.txt"] = b"".join(large)
    files["binary.bin"] = rng.randbytes(1024)
    files["empty.txt"] = b""
    return files


def _pigz_z(data, level, chunk=128 * 1024):
    body = bytearray()
    for off in range(0, max(len(data), 1), chunk):
        c = zlib.compressobj(level, zlib.DEFLATED, -15, zdict=data[max(0, off - 32768):off]) if off else zlib.compressobj(level, zlib.DEFLATED, -15)
        body += c.compress(data[off:off + chunk])
        body += c.flush(zlib.Z_FINISH if off + chunk >= len(data) else zlib.Z_SYNC_FLUSH)
    return zlib_header() + bytes(body) + zlib.adler32(data).to_bytes(4, "big")


def test_cli_commands_and_messages(engine, oracle, tmp_path):
    from zlibk_b200 import cli

    for name, data in _script_files().items():
        src, zz, back = tmp_path / name, tmp_path / (name + ".zz"), tmp_path / (name + ".out")
        src.write_bytes(data)
        for level in (None, 1, 9):
            out = io.StringIO()
            argv = ["compress", str(src), str(zz)] + ([str(level)] if level else [])
            assert cli.main(argv, out) == 0
            assert out.getvalue() == "Compressed %d bytes to %d bytes (level=%d)\n" % (len(data), os.path.getsize(zz), level or 6)
            assert zlib.decompress(zz.read_bytes()) == data  # "pigz -dz" side of the matrix
            rc, ref, _, _ = oracle.inflate(zz.read_bytes(), cap=len(data) + 64)  # the reference's own inflate
            assert rc == 1 and ref == data
            out = io.StringIO()
            assert cli.main(["inflate", str(zz), str(back)], out) == 0
            assert out.getvalue() == "Decompressed %d bytes to %d bytes\n" % (os.path.getsize(zz), len(data))
            assert back.read_bytes() == data
    out = io.StringIO()
    cli.main(["frobnicate"], out)
    assert out.getvalue().startswith("Unknown command: frobnicate\nZLib.kotlin - Pure arithmetic zlib implementation\n")
    out = io.StringIO()
    cli.main(["compress", "x"], out)
    assert out.getvalue() == "Usage: compress <input.txt> <output.zz> [level]\n"
    bad = tmp_path / "bad.zz"
    bad.write_bytes(b"\x78\x9c\x07")
    out = io.StringIO()
    cli.main(["decompress", str(bad), str(tmp_path / "bad.out")], out)
    assert out.getvalue() == "Decompression failed: -3\n"


@pytest.mark.parametrize("level", [1, 6, 9])
def test_pigz_made_streams_decode_on_the_gpu(engine, oracle, level):
    """The other direction of the matrix: `pigz -z -<level>` output (rebuilt with libz) through the GPU inflate,
    compared with the reference inflate's answer -- bytes and return code."""
    for name, data in _script_files().items():
        for stream in (zlib.compress(data, level), _pigz_z(data, level, chunk=4096), _pigz_z(data, level)):
            rc, out, used = engine.inflate(stream, len(data) + 64)
            orc, oout, _, oused = oracle.inflate(stream, cap=len(data) + 64)
            assert (rc, bytes(out), used) == (orc, oout, oused), (name, level)
            assert rc == 1 and oout == data
