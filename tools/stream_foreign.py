"""A libz-made stream through the stream handle (zk_inflate) in pieces: time, and how much input was ever pending.
usage: python tools/stream_foreign.py [mib=1024] [piece_mib=64]"""
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import zlibk_b200 as zk
from zlibk_b200 import corpus

sys.argv = sys.argv[:1] + sys.argv[1:]
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
piece = (int(sys.argv[2]) if len(sys.argv) > 2 else 64) << 20
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench import libz_stream  # noqa: E402

for kind in ("text", "mixed"):
    data = (corpus.text_like(mib << 20) if kind == "text" else corpus.mixed_entropy(mib << 20)).tobytes()
    stream = libz_stream(data, 6)
    eng = zk.default_engine()
    for mode, prefix in (("prefix decodes from 32 MiB gathered (default)", 0), ("gather everything, decode at the end", 1 << 40)):
        eng.set_option("inflate_prefix_bytes", prefix)
        p0 = eng.counter("inflate_prefix_decodes")
        z = zk.ZStream()
        z.inflateInit(15)
        got = 0
        crc = 1
        room = 256 << 20
        outbuf = bytearray(room)
        first_out_at = None
        t0 = time.perf_counter()
        pos = 0
        rc = zk.Z_OK
        while rc == zk.Z_OK:
            chunk = stream[pos: pos + piece]
            pos += len(chunk)
            last = pos >= len(stream)
            z.next_in, z.next_in_index, z.avail_in = chunk, 0, len(chunk)
            while True:
                z.next_out, z.next_out_index, z.avail_out = outbuf, 0, room
                rc = z.inflate(zk.Z_FINISH if last else zk.Z_NO_FLUSH)
                if z.next_out_index:
                    if first_out_at is None:
                        first_out_at = pos
                    crc = zlib.adler32(memoryview(outbuf)[: z.next_out_index], crc)
                    got += z.next_out_index
                if rc != zk.Z_OK or z.next_out_index < room:
                    break
            if last:
                break
        dt = time.perf_counter() - t0
        ok = rc == zk.Z_STREAM_END and got == len(data) and crc == zlib.adler32(data)
        print("%s %d MiB in %d MiB pieces, %s: %.2f s = %.2f GB/s, first output after %d MiB of input, prefix decodes %d, ok %s (rc %d, %d of %d bytes)" % (
            kind, mib, piece >> 20, mode, dt, len(data) / dt / 1e9, (first_out_at or 0) >> 20, eng.counter("inflate_prefix_decodes") - p0, ok, rc, got, len(data)), flush=True)
        z.inflateEnd()
    eng.set_option("inflate_prefix_bytes", 0)
