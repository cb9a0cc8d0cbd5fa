"""zk_inflate as a state machine (docs/API.md:239-308; SURVEY.md section 5 "checkpoint / resume"): input arrives in pieces,
every complete block is decoded once, the 32 KiB window, the unconsumed tail, the checksum and the totals are carried
from call to call.  libz's decompressobj is the oracle for the bytes and the totals (the reference has no streaming
inflate to compare with: docs only, "parity unpinned")."""
import zlib

import numpy as np
import pytest


def _feed(zk, stream, piece, wbits=15, out_room=1 << 22):
    """-> (output bytes, list of (total_in, total_out, rc) after every call)"""
    z = zk.ZStream()
    assert z.inflateInit(wbits) == zk.Z_OK
    out = bytearray()
    trace = []
    pos = 0
    while True:
        chunk = stream[pos: pos + piece]
        pos += len(chunk)
        last = pos >= len(stream)
        z.next_in, z.next_in_index, z.avail_in = chunk, 0, len(chunk)
        while True:
            z.next_out, z.next_out_index, z.avail_out = bytearray(out_room), 0, out_room
            rc = z.inflate(zk.Z_FINISH if last else zk.Z_NO_FLUSH)
            out += z.next_out[: z.next_out_index]
            trace.append((z.total_in, z.total_out, rc))
            if rc != zk.Z_OK or z.next_out_index < out_room:
                break
        if rc == zk.Z_STREAM_END or rc < 0 or last:
            break
    adler = z.adler
    z.inflateEnd()
    return bytes(out), trace, rc, adler


@pytest.mark.gpu
def test_own_stream_in_pieces_is_linear_and_incremental(engine):
    import zlibk_b200 as zk
    from zlibk_b200 import corpus

    data = corpus.mixed_entropy(64 << 20).tobytes()
    rc, stream, _ = engine.deflate(data, 6)
    stream = bytes(stream)
    eng = zk.default_engine()
    h2d0, launches0 = eng.counter("h2d_bytes"), eng.counter("kernel_launches")
    out, trace, rc, adler = _feed(zk, stream, 64 << 10)
    assert rc == zk.Z_STREAM_END and out == data and adler == zlib.adler32(data)
    calls = len(trace)
    # O(n) GPU work: every compressed byte goes up about once (plus the 32 KiB window per call), a bounded number of
    # kernel launches per call -- the round-1 implementation re-decoded the whole stream on every call
    assert eng.counter("h2d_bytes") - h2d0 <= 2 * len(stream) + calls * 40000
    assert eng.counter("kernel_launches") - launches0 <= 80 * calls
    # partial output long before the last piece, totals consistent after every call
    assert trace[len(trace) // 4][1] >= len(data) // 8
    assert all(t[0] <= len(stream) and t[1] <= len(data) for t in trace)
    assert [t[0] for t in trace] == sorted(t[0] for t in trace) and trace[-1][:2] == (len(stream), len(data))


@pytest.mark.gpu
@pytest.mark.parametrize("piece", [1 << 20, 70_001, 4096])
def test_foreign_streams_in_pieces(engine, piece):
    """libz streams have no markers: gathered, the end tried at doubling sizes (linear work), same bytes and totals."""
    import gzip

    import zlibk_b200 as zk
    from zlibk_b200 import corpus

    data = corpus.text_like(6 << 20).tobytes()
    eng = zk.default_engine()
    for stream, wbits in ((zlib.compress(data, 6), 15), (gzip.compress(data, 6, mtime=0), 31), (zlib.compress(data, 1)[2:-4], -15)):
        h2d0 = eng.counter("h2d_bytes")
        out, trace, rc, adler = _feed(zk, stream, piece, wbits)
        assert rc == zk.Z_STREAM_END and out == data
        assert trace[-1][:2] == (len(stream), len(data))
        # (gzip: the CRC-32 of the output is taken by the CRC kernel from host memory: the output goes up once more)
        assert eng.counter("h2d_bytes") - h2d0 <= 3 * len(stream) + (1 << 20) + (2 * len(data) if wbits == 31 else 0)
        if wbits == 15:
            assert adler == zlib.adler32(data)
        if wbits == 31:
            assert adler == zlib.crc32(data)
    # pigz-style framing made with libz: marked, decoded incrementally
    body = bytearray()
    for off in range(0, len(data), 128 << 10):
        c = zlib.compressobj(6, zlib.DEFLATED, -15, zdict=data[max(0, off - 32768):off]) if off else zlib.compressobj(6, zlib.DEFLATED, -15)
        body += c.compress(data[off:off + (128 << 10)])
        body += c.flush(zlib.Z_FINISH if off + (128 << 10) >= len(data) else zlib.Z_SYNC_FLUSH)
    pig = b"\x78\x9c" + bytes(body) + zlib.adler32(data).to_bytes(4, "big")
    out, trace, rc, adler = _feed(zk, pig, piece)
    assert rc == zk.Z_STREAM_END and out == data and adler == zlib.adler32(data)
    if piece >= (1 << 20):
        assert trace[0][1] > 0  # output from the first call on


@pytest.mark.gpu
def test_streaming_errors_and_trailing_bytes(engine):
    import zlibk_b200 as zk
    from zlibk_b200 import corpus

    data = corpus.text_like(1 << 20).tobytes()
    rc, own, _ = engine.deflate(data, 6)
    own = bytes(own)
    # bytes behind the stream are not consumed (zlib.h: avail_in stays)
    out, trace, rc, _ = _feed(zk, own + b"trailing garbage", 300_000)
    assert rc == zk.Z_STREAM_END and out == data and trace[-1][0] == len(own)
    # a wrong trailer: every byte comes out, then Z_DATA_ERROR (IS:419-424), and it stays the answer
    bad = bytearray(own)
    bad[-1] ^= 1
    out, trace, rc, _ = _feed(zk, bytes(bad), 100_000)
    assert rc == zk.Z_DATA_ERROR and out == data
    # a truncated stream: Z_BUF_ERROR at Z_FINISH once everything decodable has been handed out
    out, trace, rc, _ = _feed(zk, own[: len(own) // 2], 50_000)
    assert rc == zk.Z_BUF_ERROR and data.startswith(out) and len(out) > 0
    # damage inside a block: an error in the end (a flipped bit may decode to other bytes first, as with zlib: the
    # checksum is what catches it), never Z_STREAM_END
    bad = bytearray(own)
    bad[len(own) // 3] ^= 0x55
    out, trace, rc, _ = _feed(zk, bytes(bad), 64 << 10)
    assert rc in (zk.Z_DATA_ERROR, zk.Z_BUF_ERROR)
    with pytest.raises(zlib.error):
        zlib.decompress(bytes(bad))
    # bad header
    out, trace, rc, _ = _feed(zk, b"\x78\x9d" + own[2:], 4096)
    assert rc == zk.Z_DATA_ERROR and out == b""


@pytest.mark.gpu
def test_large_foreign_stream_in_pieces_fails_fast_while_unfinished(engine):
    """A libz stream has nothing to decode before its end, so the handle gathers it and tries the end when the input has
    doubled.  Every try on the unfinished stream used to cost a one-warp decode of everything gathered; now the entry-point
    search sees that only the last chunk runs out of input and says so at once (counter inflate_quick_incomplete)."""
    import time

    import zlibk_b200 as zk
    from zlibk_b200 import corpus

    data = corpus.text_like(24 << 20).tobytes()
    stream = zlib.compress(data, 6)
    eng = zk.default_engine()
    q0, f0 = eng.counter("inflate_quick_incomplete"), eng.counter("inflate_foreign_chunks")
    t0 = time.perf_counter()
    out, trace, rc, adler = _feed(zk, stream, 512 << 10, out_room=8 << 20)
    dt = time.perf_counter() - t0
    assert rc == zk.Z_STREAM_END and out == data and adler == zlib.adler32(data)
    assert eng.counter("inflate_quick_incomplete") - q0 >= 3, "the unfinished tries went through the slow decoder"
    assert eng.counter("inflate_foreign_chunks") - f0 > 100
    assert trace[-1][0] == len(stream) and trace[-1][1] == len(data)
    print("24 MiB through the handle in 512 KiB pieces: %.2f s" % dt)


@pytest.mark.gpu
@pytest.mark.parametrize("wbits", [15, 31, -15])
def test_unfinished_foreign_stream_is_decoded_up_to_its_last_entry_point(engine, wbits):
    """A libz stream fed in pieces: once enough has gathered (option inflate_prefix_bytes, 1 MiB here), everything up to the
    last entry point found in it is decoded and handed out, and the rest -- which may begin inside a byte -- is kept.  The
    bytes, totals and checksum are libz's; output must appear long before the input ends."""
    import zlibk_b200 as zk
    from zlibk_b200 import corpus

    data = corpus.mixed_entropy(20 << 20).tobytes()
    co = zlib.compressobj(6, zlib.DEFLATED, wbits)
    stream = co.compress(data) + co.flush()
    eng = zk.default_engine()
    eng.set_option("inflate_prefix_bytes", 1 << 20)
    try:
        p0 = eng.counter("inflate_prefix_decodes")
        out, trace, rc, adler = _feed(zk, stream, 700 << 10, wbits=wbits, out_room=8 << 20)
        steps = eng.counter("inflate_prefix_decodes") - p0
    finally:
        eng.set_option("inflate_prefix_bytes", 0)
    assert rc == zk.Z_STREAM_END and out == data
    assert trace[-1][0] == len(stream) and trace[-1][1] == len(data)
    if wbits == 15:
        assert adler == zlib.adler32(data)
    elif wbits == 31:
        assert adler == zlib.crc32(data)
    assert steps >= 3, steps
    early = [t for t in trace if t[0] <= len(stream) // 2]
    assert early and early[-1][1] >= len(data) // 5, "no output before half of the input had arrived"
