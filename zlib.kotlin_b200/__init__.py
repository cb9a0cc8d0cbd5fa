"""B200-native DEFLATE engine behind ZLib.kotlin's compression API.

Host side in Python (the reference's Kotlin/Native toolchain is absent in this image): a ctypes
binding of the C ABI in include/zlibk_gpu.h (`binding`), and a mirror of the reference's operator
interface (`api`: DeflateStream / InflateStream / Adler32Utils / ZLib / ZStream / ZInputStream).
There is no CPU path: without the CUDA library or a device every call raises.
"""
from .binding import (  # noqa: F401
    Z_BUF_ERROR, Z_DATA_ERROR, Z_DEFAULT_COMPRESSION, Z_DEFAULT_STRATEGY, Z_ERRNO, Z_FILTERED, Z_FINISH,
    Z_FULL_FLUSH, Z_HUFFMAN_ONLY, Z_MEM_ERROR, Z_NEED_DICT, Z_NO_FLUSH, Z_OK, Z_PARTIAL_FLUSH, Z_STREAM_END,
    Z_STREAM_ERROR, Z_SYNC_FLUSH, Engine, EngineError, library, library_path,
)
from .api import (  # noqa: F401
    Adler32, Adler32Utils, DeflateStream, InflateStream, ZInputStream, ZLib, ZStream, ZStreamException, default_engine, version,
)
