"""Inflate of streams this engine did not make (libz one-shot output, gzip members, FDICT streams): no sync markers, so
the engine searches entry points itself -- block_find_kernel accepts the bit positions where a dynamic-Huffman header that
passes every check of the decoder begins, one lane decodes each chunk between two entry points, a chunk that runs past the
next entry point (it was false) is merged with it and decoded again.  (zk_inflate.cu)

CPU part: the finder's model (tools/blockfind_model.cpp) against a plain block walk of libz streams -- every dynamic block
start is found and nothing else.  GPU part: bytes, return codes and consumed counts equal libz's / the oracle's / the
one-warp decoder's on whole, damaged, cut, dictionary and gzip streams, and the chunk count the engine reports is the one
the model predicts (i.e. the kernel accepts exactly the model's positions)."""
import ctypes
import gzip
import os
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tools", "blockfind_model.cpp")
SO = os.path.join(ROOT, "tools", "libblockfind_model.so")
SPAN = 2048  # zk_inflate.cu kFindSpan / kFindSub: one entry point per quarter span
MIN_BYTES = 256 << 10  # kForeignMinBytes


def model():
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-DBFM_LIB", "-o", SO, SRC])
    lib = ctypes.CDLL(SO)
    for f in (lib.bfm_first, lib.bfm_all, lib.bfm_walk):
        f.restype = ctypes.c_longlong
    return lib


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def walk(lib, stream, from_bit):
    b = np.frombuffer(stream, dtype=np.uint8)
    cap = len(stream) // 8 + 64
    starts = np.zeros(cap, dtype=np.uint64)
    types = np.zeros(cap, dtype=np.int32)
    k = lib.bfm_walk(_ptr(b), ctypes.c_ulonglong(len(b)), ctypes.c_ulonglong(from_bit), _ptr(starts), _ptr(types), ctypes.c_ulonglong(cap))
    assert 0 < k <= cap
    return starts[:k], types[:k]


def accepted(lib, stream, from_bit, to_bit):
    b = np.frombuffer(stream, dtype=np.uint8)
    cap = len(stream) // 8 + 64
    out = np.zeros(cap, dtype=np.uint64)
    k = lib.bfm_all(_ptr(b), ctypes.c_ulonglong(len(b)), ctypes.c_ulonglong(from_bit), ctypes.c_ulonglong(to_bit), _ptr(out), ctypes.c_ulonglong(cap))
    assert k <= cap
    return out[:k]


def first_per_span(lib, stream, min_bits):
    """the candidates block_find_kernel must report: the first accepted position behind min_bits in every span"""
    b = np.frombuffer(stream, dtype=np.uint8)
    got = []
    for s in range((len(stream) + SPAN - 1) // SPAN):
        lo = max(s * SPAN * 8, min_bits + 1)
        hi = min((s + 1) * SPAN, len(stream)) * 8
        r = lib.bfm_first(_ptr(b), ctypes.c_ulonglong(len(b)), ctypes.c_ulonglong(lo), ctypes.c_ulonglong(hi))
        if r >= 0:
            got.append(r)
    return got


def _texts():
    from zlibk_b200 import corpus

    return corpus.text_like(3 << 20).tobytes(), corpus.mixed_entropy(6 << 20).tobytes()


def test_model_finds_every_dynamic_block_and_nothing_else():
    lib = model()
    text, mixed = _texts()
    for data, level in ((text[: 1 << 20], 1), (text[: 1 << 20], 6), (text[: 1 << 20], 9), (mixed[: 3 << 20], 6)):
        stream = zlib.compress(data, level)
        starts, types = walk(lib, stream, 16)
        truth = set(int(s) for s, t in zip(starts, types) if t == 2)  # BTYPE 2, BFINAL 0
        found = set(int(x) for x in accepted(lib, stream, 0, len(stream) * 8))
        assert found == truth, (level, len(found - truth), len(truth - found))
        assert len(truth) >= 10


def test_model_rejects_noise():
    lib = model()
    noise = np.random.default_rng(5).integers(0, 256, size=1 << 20, dtype=np.uint8).tobytes()
    assert len(accepted(lib, noise, 0, len(noise) * 8)) == 0


# ---- GPU -----------------------------------------------------------------------------------------------------------


def _inflate_both(engine, stream, cap, **kw):
    """the same call with and without the entry-point search -> ((rc, bytes, consumed), (rc, bytes, consumed), chunks, merges)"""
    res = []
    counts = None
    for on in (1, 0):
        engine.set_option("inflate_foreign", on)
        c0, m0 = engine.counter("inflate_foreign_chunks"), engine.counter("inflate_foreign_merges")
        try:
            if "dictionary" in kw:
                rc, out, used = engine.inflate_dict(stream, kw["dictionary"], cap)
            elif "container" in kw:
                rc, out, used = engine.inflate_container(stream, cap, kw["container"])
            else:
                rc, out, used = engine.inflate(stream, cap)
        finally:
            engine.set_option("inflate_foreign", 1)
        res.append((rc, bytes(out), used))
        if on:
            counts = (engine.counter("inflate_foreign_chunks") - c0, engine.counter("inflate_foreign_merges") - m0)
    return res[0], res[1], counts[0], counts[1]


@pytest.mark.gpu
@pytest.mark.parametrize("level", [1, 6, 9])
def test_libz_streams_decode_through_found_entry_points(engine, level):
    import zlibk_b200 as zk

    lib = model()
    text, mixed = _texts()
    for name, data in (("text", text), ("mixed", mixed)):
        stream = zlib.compress(data, level)
        assert len(stream) >= MIN_BYTES
        fast, slow, chunks, merges = _inflate_both(engine, stream, len(data))
        assert fast == slow == (zk.Z_STREAM_END, data, len(stream)), (name, level, fast[0], slow[0], len(fast[1]))
        want = first_per_span(lib, stream, 16)
        assert chunks == len(want) + 1 and merges == 0, (name, level, chunks, len(want), merges)


@pytest.mark.gpu
def test_libz_stream_matches_oracle_inflate(engine, oracle):
    text, _ = _texts()
    stream = zlib.compress(text[: 2 << 20], 6)
    rc, out, used = engine.inflate(stream, 2 << 20)
    orc, oout, _, oused = oracle.inflate(stream, 2 << 20)
    assert (rc, bytes(out), used) == (orc, oout, oused)


@pytest.mark.gpu
def test_false_entry_points_are_merged(engine):
    """A compressed file inside the data: libz stores it (it does not compress), and its dynamic block headers sit in the
    stored payload, byte-shifted but bit-true -- entry points that are not.  The chunk in front of each runs past it, is
    merged with it and decoded again."""
    import zlibk_b200 as zk

    lib = model()
    text, _ = _texts()
    inner = zlib.compress(text[: 160 << 10], 9)[2: 40002]  # raw deflate data with a few block headers
    data = text[: 1 << 20] + inner + text[1 << 20: 2 << 20]
    stream = zlib.compress(data, 6)
    starts, types = walk(lib, stream, 16)
    truth = set(int(s) for s in starts)
    cands = first_per_span(lib, stream, 16)
    false = [c for c in cands if c not in truth]
    assert false, "the case needs at least one false entry point"
    fast, slow, chunks, merges = _inflate_both(engine, stream, len(data))
    assert fast == slow == (zk.Z_STREAM_END, data, len(stream))
    assert merges >= len(false) and chunks == len(cands) + 1 - len(false), (chunks, merges, len(cands), len(false))


@pytest.mark.gpu
def test_damaged_and_cut_streams_report_like_the_one_warp_decoder(engine, oracle):
    text, mixed = _texts()
    data = mixed[: 4 << 20]
    stream = bytearray(zlib.compress(data, 6))
    n = len(stream)
    cases = {"whole": bytes(stream), "cut in the middle": bytes(stream[: n // 2]), "cut in the trailer": bytes(stream[: n - 2]),
             "trailing bytes": bytes(stream) + b"tail"}
    for where in (n // 3, n // 2 + 12345, n - 100):
        bad = bytearray(stream)
        bad[where] ^= 0x10
        cases["bit flipped at %d" % where] = bytes(bad)
    wrong = bytearray(stream)
    wrong[-1] ^= 1
    cases["wrong adler"] = bytes(wrong)
    for name, s in cases.items():
        fast, slow, _, _ = _inflate_both(engine, s, len(data))
        assert fast == slow, (name, fast[0], slow[0], len(fast[1]), len(slow[1]), fast[2], slow[2])
        orc, oout, _, _ = oracle.inflate(s, len(data))
        assert fast[0] == orc and fast[1] == oout, (name, fast[0], orc)
    # output buffer too small: nothing is written
    fast, slow, _, _ = _inflate_both(engine, bytes(stream), len(data) - 1)
    assert fast[0] == slow[0] and fast[1] == slow[1] == b""


@pytest.mark.gpu
def test_gzip_raw_and_dictionary_streams(engine):
    import zlibk_b200 as zk

    text, mixed = _texts()
    data = text
    gz = gzip.compress(data, 6, mtime=0)
    fast, slow, chunks, _ = _inflate_both(engine, gz, len(data), container="gzip")
    assert fast == slow and fast[0] == zk.Z_STREAM_END and fast[1] == data and chunks > 8
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    raw = co.compress(mixed) + co.flush()
    fast, slow, chunks, _ = _inflate_both(engine, raw, len(mixed), container="raw")
    assert fast == slow and fast[0] == zk.Z_STREAM_END and fast[1] == mixed and chunks > 8
    dictionary = text[: 20000]
    co = zlib.compressobj(6, zlib.DEFLATED, 15, 9, zlib.Z_DEFAULT_STRATEGY, dictionary)
    body = dictionary[5000: 15000] + text[1 << 20:]
    ds = co.compress(body) + co.flush()
    fast, slow, chunks, _ = _inflate_both(engine, ds, len(body), dictionary=dictionary)
    assert fast == slow and fast[0] == zk.Z_STREAM_END and fast[1] == body and chunks > 8


@pytest.mark.gpu
def test_streams_with_stored_fixed_and_huge_blocks(engine):
    """stored blocks between Huffman blocks, a long run (one chunk decodes to megabytes), Z_FIXED blocks (no entry points:
    they ride in the chunk before them), Z_HUFFMAN_ONLY / Z_RLE streams"""
    import zlibk_b200 as zk

    rng = np.random.default_rng(3)
    text, _ = _texts()
    noise = rng.integers(0, 256, size=300000, dtype=np.uint8).tobytes()
    data = text[: 400000] + noise + bytes(5 << 20) + text[400000: 900000] + noise[:70000] + b"ab" * 600000 + text[900000: 1500000]
    for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
        co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, strategy)
        stream = co.compress(data) + co.flush()
        if strategy == zlib.Z_FIXED:  # give it a dynamic part so that there are entry points around the fixed blocks
            co2 = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_FIXED)
            fixed = co2.compress(text[: 200000]) + co2.flush(zlib.Z_FULL_FLUSH)
            co3 = zlib.compressobj(6, zlib.DEFLATED, -15)
            dyn = co3.compress(text[200000:]) + co3.flush()
            co1 = zlib.compressobj(6, zlib.DEFLATED, -15)
            head = co1.compress(text[: 200000][::-1]) + co1.flush(zlib.Z_FULL_FLUSH)
            whole = text[: 200000][::-1] + text[: 200000] + text[200000:]
            stream = b"\x78\x9c" + head + fixed + dyn + zlib.adler32(whole).to_bytes(4, "big")
            want = whole
        else:
            want = data
        if len(stream) < MIN_BYTES:
            continue
        fast, slow, chunks, _ = _inflate_both(engine, stream, len(want))
        assert fast == slow and fast[0] == zk.Z_STREAM_END and fast[1] == want, (strategy, fast[0], slow[0], len(fast[1]))


@pytest.mark.gpu
def test_reference_made_stream(engine, oracle):
    """the reference's own compressor (oracle restatement): whatever its blocks look like, both decoders agree with the
    oracle's inflate -- bytes and code"""
    text, _ = _texts()
    data = text[: 1 << 20]
    _, stream, _ = oracle.deflate(data, 6)
    fast, slow, _, _ = _inflate_both(engine, stream, len(data) + 1024)
    orc, oout, _, _ = oracle.inflate(stream, len(data) + 1024)
    assert fast == slow and fast[0] == orc and fast[1] == oout


@pytest.mark.gpu
def test_random_libz_streams(engine):
    """a few dozen streams over libz's parameter space (level, memLevel, strategy, flushes at random places, content cut from
    every corpus kind at random boundaries): the entry-point path and the one-warp path agree with libz on all of them"""
    import zlibk_b200 as zk
    from zlibk_b200 import corpus

    rng = np.random.default_rng(20261018)
    kinds = [corpus.text_like(2 << 20).tobytes(), corpus.mixed_entropy(4 << 20).tobytes(), bytes(1 << 20),
             rng.integers(0, 256, size=1 << 20, dtype=np.uint8).tobytes(), (b"0123456789abcdef" * 65536)]
    used_fast = 0
    for case in range(36):
        parts = []
        for _ in range(int(rng.integers(2, 7))):
            src = kinds[int(rng.integers(0, len(kinds)))]
            lo = int(rng.integers(0, len(src) - 1000))
            parts.append(src[lo: lo + int(rng.integers(1000, 900000))])
        data = b"".join(parts)
        level = int(rng.integers(1, 10))
        mem = int(rng.integers(1, 10))
        strategy = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_DEFAULT_STRATEGY, zlib.Z_RLE][int(rng.integers(0, 4))]
        co = zlib.compressobj(level, zlib.DEFLATED, 15, mem, strategy)
        out = []
        pos = 0
        while pos < len(data):
            step = int(rng.integers(50000, 700000))
            out.append(co.compress(data[pos: pos + step]))
            pos += step
            if rng.random() < 0.15:
                out.append(co.flush([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH, zlib.Z_PARTIAL_FLUSH if hasattr(zlib, "Z_PARTIAL_FLUSH") else zlib.Z_SYNC_FLUSH][int(rng.integers(0, 3))]))
        out.append(co.flush())
        stream = b"".join(out)
        fast, slow, chunks, _ = _inflate_both(engine, stream, len(data))
        assert fast == slow == (zk.Z_STREAM_END, data, len(stream)), (case, level, mem, strategy, fast[0], slow[0], len(fast[1]), len(data))
        used_fast += chunks > 0
    assert used_fast >= 12, used_fast
