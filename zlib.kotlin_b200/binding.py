"""ctypes binding of libzlibk_b200.so (include/zlibk_gpu.h).  Plain pointers and sizes only."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_int64, c_uint8, c_uint32, c_uint64, c_void_p

Z_OK, Z_STREAM_END, Z_NEED_DICT = 0, 1, 2
Z_ERRNO, Z_STREAM_ERROR, Z_DATA_ERROR, Z_MEM_ERROR, Z_BUF_ERROR = -1, -2, -3, -4, -5
Z_NO_COMPRESSION, Z_BEST_SPEED, Z_BEST_COMPRESSION, Z_DEFAULT_COMPRESSION = 0, 1, 9, -1
Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY = 0, 1, 2
Z_NO_FLUSH, Z_PARTIAL_FLUSH, Z_SYNC_FLUSH, Z_FULL_FLUSH, Z_FINISH = 0, 1, 2, 3, 4

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# name -> (restype, argtypes): every symbol include/zlibk_gpu.h declares
SIGNATURES = {
    "zk_version": (c_char_p, []),
    "zk_device_count": (c_int, []),
    "zk_last_error": (c_char_p, []),
    "zk_ctx_create": (c_void_p, [c_int]),
    "zk_ctx_destroy": (None, [c_void_p]),
    "zk_ctx_set_stream": (c_int, [c_void_p, c_void_p]),
    "zk_ctx_synchronize": (c_int, [c_void_p]),
    "zk_ctx_set_option": (c_int, [c_void_p, c_char_p, c_int64]),
    "zk_ctx_get_counter": (c_int64, [c_void_p, c_char_p]),
    "zk_host_alloc": (c_void_p, [c_uint64]),
    "zk_host_free": (None, [c_void_p]),
    "zk_host_register": (c_int, [c_void_p, c_uint64]),
    "zk_host_unregister": (c_int, [c_void_p]),
    "zk_adler32": (c_uint32, [c_void_p, c_uint32, c_void_p, c_uint64]),
    "zk_adler32_device": (c_int, [c_void_p, c_uint32, c_void_p, c_uint64, c_void_p]),
    "zk_adler32_combine": (c_uint32, [c_uint32, c_uint32, c_uint64]),
    "zk_compress_bound": (c_uint64, [c_uint64]),
    "zk_deflate_zlib": (c_int, [c_void_p, c_void_p, c_uint64, c_int, c_int, c_void_p, c_uint64, POINTER(c_uint64), POINTER(c_uint32)]),
    "zk_inflate_zlib": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_uint64, POINTER(c_uint64), POINTER(c_uint64)]),
    "zk_deflate_zlib_device": (c_int, [c_void_p, c_void_p, c_uint64, c_int, c_int, c_void_p, c_uint64, c_void_p]),
    "zk_inflate_zlib_device": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_uint64, c_void_p]),
    "zk_deflate_shard_device": (c_int, [c_void_p, c_void_p, c_uint32, c_void_p, c_uint64, c_int, c_int, c_void_p, c_uint64, c_void_p]),
    "zk_deflate_zlib_dict": (c_int, [c_void_p, c_void_p, c_uint32, c_void_p, c_uint64, c_int, c_int, c_void_p, c_uint64, POINTER(c_uint64), POINTER(c_uint32)]),
    "zk_inflate_zlib_dict": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_uint32, c_void_p, c_uint64, POINTER(c_uint64), POINTER(c_uint64)]),
    "zk_crc32": (c_uint32, [c_void_p, c_uint32, c_void_p, c_uint64]),
    "zk_crc32_device": (c_int, [c_void_p, c_uint32, c_void_p, c_uint64, c_void_p]),
    "zk_crc32_combine": (c_uint32, [c_uint32, c_uint32, c_uint64]),
    "zk_deflate_gzip": (c_int, [c_void_p, c_void_p, c_uint64, c_int, c_int, c_void_p, c_uint64, POINTER(c_uint64), POINTER(c_uint32)]),
    "zk_inflate_gzip": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_uint64, POINTER(c_uint64), POINTER(c_uint64)]),
    "zk_deflate_raw": (c_int, [c_void_p, c_void_p, c_uint64, c_int, c_int, c_void_p, c_uint64, POINTER(c_uint64)]),
    "zk_inflate_raw": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_uint64, POINTER(c_uint64), POINTER(c_uint64)]),
    "zk_deflate_shard": (c_int, [c_void_p, c_void_p, c_uint32, c_void_p, c_uint64, c_int, c_int, c_void_p, c_uint64, POINTER(c_uint64), POINTER(c_uint32)]),
    "zk_inflate_shard_prepare_device": (c_int, [c_void_p, c_void_p, c_uint64, c_uint32, c_uint64, c_void_p, c_void_p]),
    "zk_window_from_maps_device": (c_int, [c_void_p, c_void_p, c_uint64, c_uint32, c_void_p]),
    "zk_inflate_shard_finish_device": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "zk_zlib_header": (None, [c_int, POINTER(c_uint8)]),
    "zk_zlib_tail": (None, [c_uint32, POINTER(c_uint8)]),
    "zk_deflate_init": (c_void_p, [c_void_p, c_int, c_int]),
    "zk_deflate": (c_int, [c_void_p, c_void_p, POINTER(c_uint64), c_void_p, POINTER(c_uint64), c_int]),
    "zk_deflate_init2": (c_void_p, [c_void_p, c_int, c_int, c_int]),
    "zk_inflate_init2": (c_void_p, [c_void_p, c_int]),
    "zk_deflate_copy": (c_void_p, [c_void_p]),
    "zk_deflate_reset": (c_int, [c_void_p]),
    "zk_deflate_params": (c_int, [c_void_p, c_int, c_int]),
    "zk_deflate_set_dictionary": (c_int, [c_void_p, c_void_p, c_uint32]),
    "zk_inflate_init": (c_void_p, [c_void_p]),
    "zk_inflate": (c_int, [c_void_p, c_void_p, POINTER(c_uint64), c_void_p, POINTER(c_uint64), c_int]),
    "zk_inflate_reset": (c_int, [c_void_p]),
    "zk_inflate_set_dictionary": (c_int, [c_void_p, c_void_p, c_uint32]),
    "zk_stream_total_in": (c_uint64, [c_void_p]),
    "zk_stream_total_out": (c_uint64, [c_void_p]),
    "zk_stream_adler": (c_uint32, [c_void_p]),
    "zk_stream_end": (None, [c_void_p]),
}


class EngineError(RuntimeError):
    """The CUDA engine is missing or failed.  There is no CPU fallback."""


def library_path():
    # ZLIBK_LIB: an experimental build of the same C ABI (tools/variants.sh), never a different implementation
    return os.environ.get("ZLIBK_LIB") or os.path.join(_HERE, "libzlibk_b200.so")


def library():
    """Load libzlibk_b200.so and bind every declared symbol.  Raises if the library is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise EngineError("%s not found: build it with `python zlib.kotlin_b200/build.py` (no CPU fallback exists)" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _as_buffer(data):
    """(address, length, keepalive) of a bytes-like / numpy / torch-cpu object without copying when possible."""
    if data is None:
        return None, 0, None
    if hasattr(data, "data_ptr") and hasattr(data, "numel"):  # torch tensor on the host
        return data.data_ptr(), data.numel() * data.element_size(), data
    if hasattr(data, "ctypes") and hasattr(data, "nbytes"):  # numpy
        return data.ctypes.data, data.nbytes, data
    if isinstance(data, (bytes, bytearray, memoryview)):
        mv = memoryview(data)
        n = mv.nbytes
        if n == 0:
            return None, 0, None
        if isinstance(data, bytes):
            return ctypes.cast(ctypes.c_char_p(data), c_void_p).value, n, data
        arr = (ctypes.c_char * n).from_buffer(data if isinstance(data, bytearray) else bytearray(mv))
        return ctypes.addressof(arr), n, arr
    raise TypeError("unsupported buffer type %r" % type(data))


class Engine:
    """One zk_ctx: a CUDA device, a stream and the scratch arena.  Not thread-safe (one per thread)."""

    def __init__(self, device=0):
        self.lib = library()
        if self.lib.zk_device_count() <= 0:
            raise EngineError("no CUDA device visible: the B200 engine cannot run (there is no CPU fallback)")
        self.ctx = self.lib.zk_ctx_create(device)
        if not self.ctx:
            raise EngineError("zk_ctx_create(%d) failed: %s" % (device, self.lib.zk_last_error().decode()))
        self.device = device

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.zk_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing -------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_handle):
        return self.lib.zk_ctx_set_stream(self.ctx, cuda_stream_handle)

    def synchronize(self):
        return self.lib.zk_ctx_synchronize(self.ctx)

    def set_option(self, name, value):
        rc = self.lib.zk_ctx_set_option(self.ctx, name.encode(), int(value))
        if rc != Z_OK:
            raise ValueError("bad option %s=%r" % (name, value))

    def counter(self, name):
        return self.lib.zk_ctx_get_counter(self.ctx, name.encode())

    def last_error(self):
        return self.lib.zk_last_error().decode()

    # ---- host buffers -----------------------------------------------------------------------------
    def adler32(self, data, adler=1):
        if data is None:
            return self.lib.zk_adler32(self.ctx, adler, None, 0)
        addr, n, keep = _as_buffer(data)
        if n == 0:
            return adler & 0xFFFFFFFF
        return self.lib.zk_adler32(self.ctx, adler, addr, n)

    def compress_bound(self, n):
        return self.lib.zk_compress_bound(n)

    def deflate(self, data, level=6, strategy=Z_DEFAULT_STRATEGY, out=None):
        """-> (rc, stream bytes-like, adler).  `out` may be a preallocated numpy/torch host buffer."""
        addr, n, keep = _as_buffer(data)
        import numpy as np

        if out is None:
            out = np.empty(self.lib.zk_compress_bound(n), dtype=np.uint8)
        oaddr, ocap, okeep = _as_buffer(out)
        out_len, adler = c_uint64(0), c_uint32(0)
        rc = self.lib.zk_deflate_zlib(self.ctx, addr, n, level, strategy, oaddr, ocap, ctypes.byref(out_len), ctypes.byref(adler))
        return rc, out[: out_len.value], adler.value

    def inflate(self, stream, cap, out=None):
        """-> (rc, bytes-like out[:out_len], consumed)."""
        addr, m, keep = _as_buffer(stream)
        import numpy as np

        if out is None:
            out = np.empty(max(int(cap), 1), dtype=np.uint8)
        oaddr, ocap, okeep = _as_buffer(out)
        out_len, used = c_uint64(0), c_uint64(0)
        rc = self.lib.zk_inflate_zlib(self.ctx, addr, m, oaddr, min(int(cap), ocap), ctypes.byref(out_len), ctypes.byref(used))
        return rc, out[: out_len.value], used.value

    # ---- device buffers (torch tensors on this device; asynchronous on the context stream) -------------
    def deflate_device(self, d_in, n, level, strategy, d_out, cap, d_result):
        return self.lib.zk_deflate_zlib_device(self.ctx, d_in, n, level, strategy, d_out, cap, d_result)

    def inflate_device(self, d_in, m, d_out, cap, d_result):
        return self.lib.zk_inflate_zlib_device(self.ctx, d_in, m, d_out, cap, d_result)

    def adler32_device(self, d_buf, n, d_out, adler=1):
        return self.lib.zk_adler32_device(self.ctx, adler, d_buf, n, d_out)

    def deflate_shard_device(self, d_hist, hist_len, d_in, n, level, strategy, d_body, cap, d_meta):
        return self.lib.zk_deflate_shard_device(self.ctx, d_hist, hist_len, d_in, n, level, strategy, d_body, cap, d_meta)

    def crc32(self, data, crc=0):
        addr, n, keep = _as_buffer(data)
        return self.lib.zk_crc32(self.ctx, crc, addr, n) if n else crc

    def crc32_combine(self, crc1, crc2, len2):
        return self.lib.zk_crc32_combine(crc1, crc2, len2)

    def deflate_container(self, data, level=6, container="gzip", strategy=Z_DEFAULT_STRATEGY):
        """-> (rc, stream bytes-like, crc32 or None) as a gzip member or a raw deflate stream."""
        import numpy as np

        addr, n, keep = _as_buffer(data)
        out = np.empty(self.lib.zk_compress_bound(n) + 32, dtype=np.uint8)
        oaddr, ocap, okeep = _as_buffer(out)
        out_len, crc = c_uint64(0), c_uint32(0)
        if container == "gzip":
            rc = self.lib.zk_deflate_gzip(self.ctx, addr, n, level, strategy, oaddr, ocap, ctypes.byref(out_len), ctypes.byref(crc))
            return rc, out[: out_len.value], crc.value
        rc = self.lib.zk_deflate_raw(self.ctx, addr, n, level, strategy, oaddr, ocap, ctypes.byref(out_len))
        return rc, out[: out_len.value], None

    def inflate_container(self, stream, cap, container="gzip"):
        """-> (rc, bytes-like out[:out_len], consumed)."""
        import numpy as np

        addr, m, keep = _as_buffer(stream)
        out = np.empty(max(int(cap), 1), dtype=np.uint8)
        oaddr, ocap, okeep = _as_buffer(out)
        out_len, used = c_uint64(0), c_uint64(0)
        fn = self.lib.zk_inflate_gzip if container == "gzip" else self.lib.zk_inflate_raw
        rc = fn(self.ctx, addr, m, oaddr, min(int(cap), ocap), ctypes.byref(out_len), ctypes.byref(used))
        return rc, out[: out_len.value], used.value

    def deflate_dict(self, dictionary, data, level=6, strategy=Z_DEFAULT_STRATEGY):
        """zlib stream with FDICT: -> (rc, stream bytes-like, adler of the data)."""
        import numpy as np

        addr, n, keep = _as_buffer(data)
        daddr, dn, dkeep = _as_buffer(dictionary)
        out = np.empty(self.lib.zk_compress_bound(n) + 16, dtype=np.uint8)
        oaddr, ocap, okeep = _as_buffer(out)
        out_len, adler = c_uint64(0), c_uint32(0)
        rc = self.lib.zk_deflate_zlib_dict(self.ctx, daddr, dn, addr, n, level, strategy, oaddr, ocap, ctypes.byref(out_len), ctypes.byref(adler))
        return rc, out[: out_len.value], adler.value

    def inflate_dict(self, stream, dictionary, cap):
        """-> (rc, bytes-like out[:out_len], consumed) for a stream whose header has FDICT."""
        import numpy as np

        addr, m, keep = _as_buffer(stream)
        daddr, dn, dkeep = _as_buffer(dictionary)
        out = np.empty(max(int(cap), 1), dtype=np.uint8)
        oaddr, ocap, okeep = _as_buffer(out)
        out_len, used = c_uint64(0), c_uint64(0)
        rc = self.lib.zk_inflate_zlib_dict(self.ctx, addr, m, daddr, dn, oaddr, min(int(cap), ocap), ctypes.byref(out_len), ctypes.byref(used))
        return rc, out[: out_len.value], used.value

    def deflate_shard(self, hist, data, level=6, strategy=Z_DEFAULT_STRATEGY, out=None):
        """Host buffers: -> (rc, body bytes-like, adler of the shard).  `hist` = up to 32 KiB before the shard or None."""
        import numpy as np

        addr, n, keep = _as_buffer(data)
        haddr, hn, hkeep = _as_buffer(hist) if hist is not None and len(hist) else (None, 0, None)
        if out is None:
            out = np.empty(self.lib.zk_compress_bound(n), dtype=np.uint8)
        oaddr, ocap, okeep = _as_buffer(out)
        out_len, adler = c_uint64(0), c_uint32(0)
        rc = self.lib.zk_deflate_shard(self.ctx, haddr, hn, addr, n, level, strategy, oaddr, ocap, ctypes.byref(out_len), ctypes.byref(adler))
        return rc, out[: out_len.value], adler.value

    def inflate_shard_prepare_device(self, d_body, m, hist_len, cap, d_tail_map, d_result):
        return self.lib.zk_inflate_shard_prepare_device(self.ctx, d_body, m, hist_len, cap, d_tail_map, d_result)

    def window_from_maps_device(self, d_maps, stride_bytes, count, d_window):
        return self.lib.zk_window_from_maps_device(self.ctx, d_maps, stride_bytes, count, d_window)

    def inflate_shard_finish_device(self, d_entry_window, d_out, d_result):
        return self.lib.zk_inflate_shard_finish_device(self.ctx, d_entry_window, d_out, d_result)

    def adler32_combine(self, a1, a2, len2):
        return self.lib.zk_adler32_combine(a1, a2, len2)

    def zlib_header(self, level):
        buf = (c_uint8 * 2)()
        self.lib.zk_zlib_header(level, buf)
        return bytes(buf)

    def zlib_tail(self, adler):
        buf = (c_uint8 * 6)()
        self.lib.zk_zlib_tail(adler, buf)
        return bytes(buf)
