"""CPU tests: the C-ABI library loads, exports every symbol include/zlibk_gpu.h declares, its pure-host
helpers are right, and the product path fails loudly without a device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import zlib

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import zlibk_b200

    if not os.path.exists(zlibk_b200.library_path()):
        import __graft_entry__

        __graft_entry__.build()
    return zlibk_b200.library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "zlibk_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zk_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from zlibk_b200 import binding

    names = declared_symbols()
    assert len(names) >= 30
    out = subprocess.check_output(["nm", "-D", "--defined-only", binding.library_path()], text=True)
    exported = set(re.findall(r" T (zk_[a-z0-9_]+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)
    assert set(names) == set(binding.SIGNATURES), sorted(set(names) ^ set(binding.SIGNATURES))


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "zlibk_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # signatures only, not the prose
    assert "torch" not in text and "at::" not in text and "#include <cuda" not in text


def test_library_is_sm100a(lib):
    from zlibk_b200 import binding

    out = subprocess.run(["cuobjdump", "-lelf", binding.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_version(lib):
    assert lib.zk_version().decode().startswith("1.0.2")  # common/Constants.kt:37-44


def test_adler_combine_matches_libz(lib):
    import random

    rng = random.Random(5)
    for _ in range(200):
        a = rng.randbytes(rng.choice([0, 1, 5, 5552, 70000]))
        b = rng.randbytes(rng.choice([0, 1, 7, 65521, 100000]))
        got = lib.zk_adler32_combine(zlib.adler32(a), zlib.adler32(b), len(b))
        assert got == zlib.adler32(a + b)


def test_compress_bound_covers_stored_fallback(lib):
    for n in (0, 1, 32768, 65535, 1 << 20, 1 << 30, 8 << 30):
        blocks = n // 32768 + 1
        assert lib.zk_compress_bound(n) >= n + 10 * blocks + 8
        assert lib.zk_compress_bound(n) >= n + (n >> 12) + (n >> 14) + 11  # docs/API.md:690


@pytest.mark.parametrize("level,flg", [(-1, 0x01), (0, 0x01), (1, 0x01), (2, 0x5E), (5, 0x5E), (6, 0x9C), (8, 0x9C), (9, 0xDA)])
def test_header_bytes(lib, level, flg):
    buf = (ctypes.c_uint8 * 2)()
    lib.zk_zlib_header(level, buf)
    assert bytes(buf) == bytes([0x78, flg])  # DeflateStream.kt:27-50


def test_tail_is_a_valid_stream_end(lib):
    buf = (ctypes.c_uint8 * 6)()
    lib.zk_zlib_tail(1, buf)
    assert zlib.decompress(bytes([0x78, 0x9C]) + bytes(buf)) == b""


def test_null_context_is_rejected(lib):
    n = ctypes.c_uint64(0)
    assert lib.zk_deflate_zlib(None, None, 0, 6, 0, None, 0, ctypes.byref(n), None) == -2
    assert lib.zk_inflate_zlib(None, None, 0, None, 0, ctypes.byref(n), None) == -2
    assert lib.zk_adler32(None, 7, None, 0) == 1  # null buffer -> 1 whatever the seed (AU:37)


def test_level_above_nine_is_stream_error_without_a_device():
    import io

    import zlibk_b200

    sink = io.BytesIO()
    assert zlibk_b200.DeflateStream.compressZlibResult(io.BytesIO(b""), sink, level=10) == (zlibk_b200.Z_STREAM_ERROR, 0)
    assert sink.getvalue() == b""  # DeflateParameterValidationTest.kt:12-20


def test_product_path_fails_loudly_without_gpu(lib):
    import zlibk_b200

    if lib.zk_device_count() > 0:
        pytest.skip("a device is present")
    with pytest.raises(zlibk_b200.EngineError):
        zlibk_b200.Engine(0)
    with pytest.raises(zlibk_b200.EngineError):
        zlibk_b200.ZLib.compress(b"abc")


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "zlib.kotlin_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".kt", ".def")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libzk_oracle" not in text and "oracle_binding" not in text and "zko_" not in text, f
