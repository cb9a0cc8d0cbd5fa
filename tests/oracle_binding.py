"""ctypes binding of oracle/libzk_oracle.so -- the CPU restatement of the reference.

Test infrastructure: imported only by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.
"""
import ctypes
from ctypes import POINTER, c_char_p, c_int, c_size_t, c_uint32, c_uint64, c_void_p

Z_OK, Z_STREAM_END, Z_NEED_DICT = 0, 1, 2
Z_ERRNO, Z_STREAM_ERROR, Z_DATA_ERROR, Z_MEM_ERROR, Z_BUF_ERROR = -1, -2, -3, -4, -5
ESCAPED_EXCEPTION = -100


class Oracle:
    def __init__(self, path):
        L = ctypes.CDLL(path)
        L.zko_adler32.restype = c_uint32
        L.zko_adler32.argtypes = [c_uint32, c_void_p, c_size_t]
        L.zko_inflate_zlib.restype = c_int
        L.zko_inflate_zlib.argtypes = [c_char_p, c_size_t, c_void_p, c_size_t, POINTER(c_size_t), POINTER(c_uint64), POINTER(c_size_t)]
        L.zko_deflate_zlib.restype = c_int
        L.zko_deflate_zlib.argtypes = [c_char_p, c_size_t, c_int, POINTER(c_void_p), POINTER(c_size_t), POINTER(c_uint64)]
        L.zko_free.argtypes = [c_void_p]
        L.zko_build_lengths.argtypes = [POINTER(c_int), c_int, c_int, c_int, POINTER(c_int)]
        L.zko_build_encoder.argtypes = [POINTER(c_int), c_int, POINTER(c_int), POINTER(c_int)]
        L.zko_build_full.restype = c_int
        L.zko_build_full.argtypes = [POINTER(c_int), c_int, POINTER(c_int), POINTER(c_int)]
        L.zko_version.restype = c_char_p
        self.L = L

    def version(self):
        return self.L.zko_version().decode()

    def adler32(self, data, adler=1):
        """Adler32Utils.adler32(adler, buf, 0, len); data=None models a null buffer."""
        if data is None:
            return self.L.zko_adler32(adler, None, 0)
        data = bytes(data)
        buf = ctypes.create_string_buffer(data, len(data)) if data else ctypes.create_string_buffer(1)
        return self.L.zko_adler32(adler, ctypes.cast(buf, c_void_p), len(data))

    def inflate(self, stream, cap=None):
        """InflateStream.inflateZlib -> (rc, sink_bytes, pair_second, consumed)."""
        stream = bytes(stream)
        cap = cap if cap is not None else max(4096, len(stream) * 1040)
        out = ctypes.create_string_buffer(cap if cap else 1)
        n, rep, used = c_size_t(), c_uint64(), c_size_t()
        rc = self.L.zko_inflate_zlib(stream, len(stream), ctypes.cast(out, c_void_p), cap, n, rep, used)
        return rc, out.raw[: n.value], rep.value, used.value

    def deflate(self, data, level=6):
        """DeflateStream.compressZlibResult -> (rc, stream_bytes, bytes_in)."""
        data = bytes(data)
        p, n, used = c_void_p(), c_size_t(), c_uint64()
        rc = self.L.zko_deflate_zlib(data, len(data), level, p, n, used)
        out = ctypes.string_at(p, n.value) if p.value else b""
        self.L.zko_free(p)
        return rc, out, used.value

    def build_lengths(self, freq, max_bits, ensure=-1):
        n = len(freq)
        f = (c_int * n)(*freq)
        o = (c_int * n)()
        self.L.zko_build_lengths(f, n, max_bits, ensure, o)
        return list(o)

    def build_encoder(self, lengths):
        n = len(lengths)
        l = (c_int * n)(*lengths)
        c, b = (c_int * n)(), (c_int * n)()
        self.L.zko_build_encoder(l, n, c, b)
        return list(c), list(b)

    def build_full(self, lengths):
        n = len(lengths)
        l = (c_int * n)(*lengths)
        bits, vals = (c_int * 32768)(), (c_int * 32768)()
        m = self.L.zko_build_full(l, n, bits, vals)
        if m < 0:
            return m, [], []
        return m, list(bits[: 1 << m]), list(vals[: 1 << m])
