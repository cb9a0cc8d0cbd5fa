"""Host-side mirror of the reference's operator interface for the compression path.

The reference is Kotlin; its toolchain is not in this image, so the host side above the C ABI is
Python here and Kotlin in `kotlin/` (unbuilt).  Names, argument meaning and return values follow
the reference so that the parity tests read like the reference's own tests:

  DeflateStream.compressZlib / compressZlibResult   deflate/DeflateStream.kt:302-328
  InflateStream.inflateZlib                         inflate/InflateStream.kt:376-438
  Adler32Utils.adler32, Adler32().adler32           bitwise/checksum/Adler32Utils.kt:31-57, deflate/Adler32.kt:35
  version()                                         common/Constants.kt:44
  ZLib / ZStream / ZInputStream / ZStreamException  documented-only surface, docs/API.md:170-735, README.md:84-92

Okio's BufferedSource / BufferedSink are played by any object with read(n) / write(bytes)
(io.BytesIO, open files).  Every call runs on the GPU engine; there is no CPU path.
"""
import io
import threading

from . import binding as _b

_tls = threading.local()


def default_engine():
    """One Engine per host thread ("one ZStream per thread", docs/API.md:725)."""
    eng = getattr(_tls, "engine", None)
    if eng is None:
        eng = _b.Engine(0)
        _tls.engine = eng
    return eng


def version():
    return _b.library().zk_version().decode()


class ZStreamException(Exception):
    """docs/API.md:310-318"""


def _drain(source):
    if isinstance(source, (bytes, bytearray, memoryview)):
        return bytes(source)
    data = source.read()
    return data if data is not None else b""


class DeflateStream:
    """object DeflateStream (deflate/DeflateStream.kt:23)"""

    @staticmethod
    def compressZlib(source, sink, level=6):
        """Returns the bytes consumed, like DS:302-311.  level <= 0 stored, 1 fastest, >= 2 dynamic."""
        rc, used = DeflateStream.compressZlibResult(source, sink, level)
        if rc != _b.Z_OK:
            raise ZStreamException("compressZlib failed: %d" % rc)
        return used

    @staticmethod
    def compressZlibResult(source, sink, level=6):
        """(Z_OK, bytesIn), or (Z_STREAM_ERROR, 0) with nothing written for level > 9 (DS:318-328)."""
        if level > 9:
            return _b.Z_STREAM_ERROR, 0
        data = _drain(source)
        rc, out, _ = default_engine().deflate(data, level)
        if rc != _b.Z_OK:
            return rc, 0
        sink.write(bytes(out))
        if hasattr(sink, "flush"):
            sink.flush()
        return _b.Z_OK, len(data)


class InflateStream:
    """object InflateStream (inflate/InflateStream.kt:25)"""

    @staticmethod
    def inflateZlib(source, sink, max_output=None):
        """(rc, bytesOut); rc is Z_STREAM_END on success (IS:424).  Bytes decoded before an error are
        still written to the sink, as the reference does."""
        data = _drain(source)
        eng = default_engine()
        cap = max_output if max_output is not None else max(65536, len(data) * 8)
        while True:
            rc, out, used = eng.inflate(data, cap)
            if rc == _b.Z_BUF_ERROR and len(out) == 0 and max_output is None and cap < len(data) * 1100 + 65536:
                cap *= 8  # the output did not fit (or the input is truncated): retry up to DEFLATE's 1032:1
                continue
            break
        if len(out):
            sink.write(bytes(out))
        if hasattr(source, "seek") and hasattr(source, "tell") and not isinstance(source, (bytes, bytearray)):
            try:  # leave a second member unread, as the reference leaves it in the source
                source.seek(source.tell() - (len(data) - used))
            except (OSError, ValueError):
                pass
        return rc, len(out)


class Adler32Utils:
    """object Adler32Utils (bitwise/checksum/Adler32Utils.kt:20)"""

    @staticmethod
    def adler32(adler, buf, index, length):
        if buf is None:
            return 1  # AU:37
        view = memoryview(buf)[index:index + length]
        return default_engine().adler32(bytes(view), adler & 0xFFFFFFFF)


class Adler32:
    """class Adler32 (deflate/Adler32.kt:25-43)"""

    def adler32(self, adler, buf, index, length):
        return Adler32Utils.adler32(adler, buf, index, length)


class ZLib:
    """ZLib.compress / ZLib.decompress (README.md:84-92)."""

    @staticmethod
    def compress(data, level=6):
        rc, out, _ = default_engine().deflate(bytes(data), level)
        if rc != _b.Z_OK:
            raise ZStreamException("compress failed: %d" % rc)
        return bytes(out)

    @staticmethod
    def decompress(data):
        sink = io.BytesIO()
        rc, _ = InflateStream.inflateZlib(bytes(data), sink)
        if rc != _b.Z_STREAM_END:
            raise ZStreamException("decompress failed: %d" % rc)
        return sink.getvalue()


class ZStream:
    """The documented ZStream (docs/API.md:205-281): caller-managed next_in/next_out windows over the
    streaming handles of the C ABI.  Flush semantics follow zlib.h (the reference defines the constants
    only, common/Constants.kt:79-83)."""

    def __init__(self):
        self.next_in = b""
        self.next_in_index = 0
        self.avail_in = 0
        self.total_in = 0
        self.next_out = bytearray()
        self.next_out_index = 0
        self.avail_out = 0
        self.total_out = 0
        self.msg = None
        self.adler = 1
        self.data_type = 2  # Z_UNKNOWN
        self._h = None
        self._mode = None
        self._lib = _b.library()

    # -- deflate ---------------------------------------------------------------------------------
    @staticmethod
    def _container(windowBits):
        """zlib.h: 8..15 zlib, -8..-15 raw deflate, 24..31 gzip -> container id, or None."""
        if 8 <= windowBits <= 15:
            return 0
        if -15 <= windowBits <= -8:
            return 1
        if 24 <= windowBits <= 31:
            return 2
        return None

    def deflateInit(self, level=6, method=8, windowBits=15, memLevel=8, strategy=_b.Z_DEFAULT_STRATEGY):
        container = self._container(windowBits)
        if method != 8 or container is None or level > 9:
            return _b.Z_STREAM_ERROR
        self._end()
        self._h = self._lib.zk_deflate_init2(default_engine().ctx, level, strategy, container)
        self._mode = "deflate"
        return _b.Z_OK if self._h else _b.Z_STREAM_ERROR

    def deflateCopy(self, dest):
        """docs/API.md:253: dest becomes an independent copy of this deflate stream."""
        if self._mode != "deflate" or not self._h or not isinstance(dest, ZStream):
            return _b.Z_STREAM_ERROR
        dest._end()
        dest._h = self._lib.zk_deflate_copy(self._h)
        dest._mode = "deflate"
        for f in ("next_in", "next_in_index", "avail_in", "total_in", "next_out_index", "avail_out", "total_out", "adler", "data_type", "msg"):
            setattr(dest, f, getattr(self, f))
        dest.next_out = bytearray(self.next_out)
        return _b.Z_OK if dest._h else _b.Z_MEM_ERROR

    def deflate(self, flush=_b.Z_NO_FLUSH):
        return self._run(self._lib.zk_deflate, "deflate", flush)

    def deflateParams(self, level, strategy):
        return self._lib.zk_deflate_params(self._h, level, strategy) if self._mode == "deflate" else _b.Z_STREAM_ERROR

    def deflateSetDictionary(self, dictionary, dictLength=None):
        n = len(dictionary) if dictLength is None else dictLength
        return self._lib.zk_deflate_set_dictionary(self._h, bytes(dictionary[:n]), n) if self._mode == "deflate" else _b.Z_STREAM_ERROR

    def deflateReset(self):
        self.total_in = self.total_out = 0
        return self._lib.zk_deflate_reset(self._h) if self._mode == "deflate" else _b.Z_STREAM_ERROR

    def deflateEnd(self):
        return self._end()

    # -- inflate ---------------------------------------------------------------------------------
    def inflateInit(self, windowBits=15):
        container = self._container(windowBits)
        if container is None:
            return _b.Z_STREAM_ERROR
        self._end()
        self._h = self._lib.zk_inflate_init2(default_engine().ctx, container)
        self._mode = "inflate"
        return _b.Z_OK if self._h else _b.Z_STREAM_ERROR

    def inflate(self, flush=_b.Z_NO_FLUSH):
        return self._run(self._lib.zk_inflate, "inflate", flush)

    def inflateSetDictionary(self, dictionary, dictLength=None):
        n = len(dictionary) if dictLength is None else dictLength
        return self._lib.zk_inflate_set_dictionary(self._h, bytes(dictionary[:n]), n) if self._mode == "inflate" else _b.Z_STREAM_ERROR

    def inflateSync(self):
        return _b.Z_DATA_ERROR  # no sync point search: the reference has none either (docs only)

    def inflateReset(self):
        self.total_in = self.total_out = 0
        return self._lib.zk_inflate_reset(self._h) if self._mode == "inflate" else _b.Z_STREAM_ERROR

    def inflateEnd(self):
        return self._end()

    def free(self):
        self._end()
        self.next_in = b""
        self.next_out = bytearray()

    # -- shared ----------------------------------------------------------------------------------
    def _run(self, fn, mode, flush):
        import ctypes

        if self._mode != mode or not self._h:
            return _b.Z_STREAM_ERROR
        src = bytes(memoryview(self.next_in)[self.next_in_index:self.next_in_index + self.avail_in])
        in_len = ctypes.c_uint64(len(src))
        out_len = ctypes.c_uint64(self.avail_out)
        # the engine writes straight into next_out when that is a bytearray with the room it claims (no staging copy)
        direct = isinstance(self.next_out, bytearray) and self.avail_out > 0 and len(self.next_out) >= self.next_out_index + self.avail_out
        if direct:
            out_buf = (ctypes.c_uint8 * self.avail_out).from_buffer(self.next_out, self.next_out_index)
        else:
            out_buf = (ctypes.c_uint8 * max(self.avail_out, 1))()
        rc = fn(self._h, src if src else None, ctypes.byref(in_len), out_buf, ctypes.byref(out_len), flush)
        used, made = in_len.value, out_len.value
        self.next_in_index += used
        self.avail_in -= used
        if made and not direct:
            self.next_out[self.next_out_index:self.next_out_index + made] = ctypes.string_at(out_buf, made)
        del out_buf
        self.next_out_index += made
        self.avail_out -= made
        self.total_in = self._lib.zk_stream_total_in(self._h)
        self.total_out = self._lib.zk_stream_total_out(self._h)
        self.adler = self._lib.zk_stream_adler(self._h)
        if rc < 0:
            self.msg = "zlibk error %d" % rc
        return rc

    def _end(self):
        if self._h:
            self._lib.zk_stream_end(self._h)
            self._h = None
            self._mode = None
            return _b.Z_OK
        return _b.Z_STREAM_ERROR


class ZInputStream:
    """ZInputStream(InputStream) (docs/API.md:283-308): read() yields decompressed bytes."""

    def __init__(self, stream):
        self._src = stream
        self._z = ZStream()
        if self._z.inflateInit() != _b.Z_OK:
            raise ZStreamException("inflateInit failed")
        self._buf = b""
        self._eof = False

    def _fill(self):
        if self._eof:
            return
        data = self._src.read()
        z = self._z
        z.next_in, z.next_in_index, z.avail_in = data or b"", 0, len(data or b"")
        pieces = []
        while True:
            z.next_out, z.next_out_index, z.avail_out = bytearray(1 << 20), 0, 1 << 20
            rc = z.inflate(_b.Z_FINISH)
            pieces.append(bytes(z.next_out[:z.next_out_index]))
            if rc == _b.Z_STREAM_END:
                break
            if rc not in (_b.Z_OK,):
                raise ZStreamException("inflate failed: %d" % rc)
        self._buf = b"".join(pieces)
        self._eof = True

    def read(self, n=-1):
        self._fill()
        if n is None or n < 0:
            out, self._buf = self._buf, b""
        else:
            out, self._buf = self._buf[:n], self._buf[n:]
        return out

    def available(self):
        self._fill()
        return len(self._buf)

    def getTotalIn(self):
        return self._z.total_in

    def getTotalOut(self):
        return self._z.total_out

    def close(self):
        self._z.inflateEnd()
        if hasattr(self._src, "close"):
            self._src.close()
