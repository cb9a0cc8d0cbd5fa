"""lz77_kernel against its CPU model (tools/lzmodel2.cpp), token for token.

The model restates the kernel's algorithm decision for decision (exact hash chains with 16-bit links, the best match of
every position, one speculative parse per 16-byte segment, the stitch, the real prefix, the adaptive probe rule), checks
that its own parse reproduces the input, and is what the ratio parameters were tuned with.  Equality of the two token streams means the
kernel's result is a function of the data alone (no dependence on warp timing) and is the parse that was designed.
Parity of the STREAM with the reference is tests/test_gpu_parity.py's business; this file pins the tokeniser.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODEL_SRC = os.path.join(ROOT, "tools", "lzmodel2.cpp")
MODEL_SO = os.path.join(ROOT, "tools", "liblzmodel2.so")

# zlib.kotlin_b200/csrc/zk_lz77.cu level_params(): chain, lazy, max_lazy, probe_adaptive, lazy2, far4
LEVELS = {
    1: (1, 0, 16, 1, 0, 1024),
    2: (1, 1, 4, 1, 0, 1024),
    3: (1, 1, 8, 1, 0, 1024),
    4: (1, 1, 16, 1, 0, 1024),
    5: (1, 1, 16, 1, 8, 1024),
    6: (1, 1, 16, 1, 16, 512),
    7: (2, 1, 16, 1, 16, 512),
    8: (4, 1, 32, 1, 16, 1024),
    9: (8, 1, 258, 0, 16, 1024),
}


def build_model():
    if not os.path.exists(MODEL_SO) or os.path.getmtime(MODEL_SO) < max(
        os.path.getmtime(MODEL_SRC), os.path.getmtime(os.path.join(ROOT, "tools", "lzcost.hpp"))
    ):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-DLZM2_LIB", "-o", MODEL_SO, MODEL_SRC])
    lib = ctypes.CDLL(MODEL_SO)
    lib.lzm2_tokens.restype = ctypes.c_longlong
    lib.lzm2_tokens.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_void_p, ctypes.c_int,
                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return lib


def model_tokens(lib, data, level, unit, block=32768, strategy=0):
    chain, lazy, max_lazy, adaptive, lazy2, far4 = LEVELS[level]
    min_len, filtered = 3, 0
    if strategy == 1:
        filtered = 1
    if strategy == 2:
        chain, min_len = 0, 259
    params = np.array([chain, 128, lazy, max_lazy, 256, far4, min_len, filtered, adaptive, 4, 8, 16, block, unit, lazy2], dtype=np.int32)
    n = len(data)
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    buf = np.ascontiguousarray(buf)
    tokens = np.zeros(max(n, 1), dtype=np.uint32)
    nb_cap = n // block + n // unit + 4
    bb = np.zeros(nb_cap, dtype=np.uint32)
    bc = np.zeros(nb_cap, dtype=np.uint32)
    nblocks = ctypes.c_ulonglong(0)
    nt = lib.lzm2_tokens(buf.ctypes.data, n, 0, params.ctypes.data, len(params), tokens.ctypes.data, bb.ctypes.data, bc.ctypes.data,
                         ctypes.byref(nblocks))
    assert nt >= 0, "the model's own parse does not reproduce the input"
    nb = nblocks.value
    return [tokens[bb[i]: bb[i] + bc[i]] for i in range(nb)]


def test_model_builds_and_parses():
    """CPU: the model compiles, and its stitched parse reproduces the input on a few shapes (lzm2_tokens verifies)."""
    lib = build_model()
    rng = np.random.default_rng(7)
    words = [bytes(rng.integers(97, 123, size=int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(300)]
    text = b" ".join(words[int(i)] for i in rng.integers(0, 300, size=40000))
    cases = [b"", b"a", b"abcabcabcabcabcabc" * 50, bytes(70000), text[:150000], rng.integers(0, 256, size=100000, dtype=np.uint8).tobytes(),
             np.arange(60000, dtype=np.uint32).tobytes()]
    for data in cases:
        for level in (1, 6, 9):
            blocks = model_tokens(lib, data, level, unit=65536)
            covered = sum(int(((b >> 16) == 0).sum()) + int((b >> 16).sum()) for b in blocks)
            assert covered == len(data)


def gpu_tokens(engine, data, level, unit, strategy=0):
    import torch

    n = len(data)
    dev = torch.device("cuda", engine.device)
    host = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    d_in = torch.from_numpy(np.ascontiguousarray(host).copy()).to(dev)
    cap = engine.compress_bound(n)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(4, dtype=torch.int64, device=dev)
    nb = (n + 32767) // 32768
    d_tok = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    d_blk = torch.zeros(2 * max(nb, 1), dtype=torch.int32, device=dev)
    engine.set_option("chunk_bytes", unit)
    engine.set_option("lz_dump_tokens", d_tok.data_ptr())
    engine.set_option("lz_dump_blocks", d_blk.data_ptr())
    try:
        rc = engine.deflate_device(d_in.data_ptr(), n, level, strategy, d_out.data_ptr(), cap, d_res.data_ptr())
        engine.synchronize()
    finally:
        engine.set_option("lz_dump_tokens", 0)
        engine.set_option("lz_dump_blocks", 0)
        engine.set_option("chunk_bytes", 0)
    assert rc == 0
    res = d_res.cpu().tolist()
    assert res[2] == 0
    tok = d_tok.cpu().numpy().view(np.uint32)
    blk = d_blk.cpu().numpy().view(np.uint32)
    return [tok[blk[i]: blk[i] + blk[nb + i]] for i in range(nb)], bytes(d_out[: res[0]].cpu().numpy())


def corpus_cases():
    import zlibk_b200
    from zlibk_b200 import corpus

    rng = np.random.default_rng(11)
    yield "text", corpus.text_like(3 << 20)
    yield "mixed", corpus.mixed_entropy(6 << 20)
    yield "random", rng.integers(0, 256, size=(1 << 20) + 12345, dtype=np.uint8)
    yield "zeros", np.zeros(700001, dtype=np.uint8)
    yield "counters", np.arange(300000, dtype=np.uint32).view(np.uint8)
    yield "short", np.frombuffer(b"abcabcabc\0\0\0hello hello hello", dtype=np.uint8)
    yield "ragged", corpus.text_like(1 << 20)[: 65536 * 3 + 4097]


@pytest.mark.gpu
@pytest.mark.parametrize("level", [1, 3, 6, 7, 9])
def test_kernel_tokens_equal_model(engine, level):
    import zlib

    lib = build_model()
    for name, data in corpus_cases():
        for unit in (65536, 524288):
            want = model_tokens(lib, data, level, unit)
            got, stream = gpu_tokens(engine, data, level, unit)
            assert zlib.decompress(stream) == data.tobytes(), (name, level, unit)
            assert len(got) == len(want), (name, level, unit)
            for b, (g, w) in enumerate(zip(got, want)):
                if len(g) != len(w) or not np.array_equal(g, w):
                    k = int(np.argmax(g[: min(len(g), len(w))] != w[: min(len(g), len(w))])) if len(g) and len(w) else 0
                    raise AssertionError("%s level %d unit %d block %d: %d vs %d tokens, first difference at token %d: gpu %s model %s" % (
                        name, level, unit, b, len(g), len(w), k, [hex(int(x)) for x in g[k: k + 4]], [hex(int(x)) for x in w[k: k + 4]]))


@pytest.mark.gpu
def test_kernel_tokens_strategies(engine):
    lib = build_model()
    from zlibk_b200 import corpus

    data = corpus.mixed_entropy(2 << 20)
    for strategy in (1, 2):
        want = model_tokens(lib, data, 6, 131072, strategy=strategy)
        got, _ = gpu_tokens(engine, data, 6, 131072, strategy=strategy)
        assert all(np.array_equal(g, w) for g, w in zip(got, want))
