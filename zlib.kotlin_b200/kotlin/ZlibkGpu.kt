// ZlibkGpu.kt -- Kotlin/Native host side of the B200 DEFLATE engine.
//
// UNBUILT in this repository's image (no Kotlin/Native toolchain).  It is the thin pass-through a maintainer
// adds to ZLib.kotlin's nativeMain source set: same signatures as the reference's entry points
// (deflate/DeflateStream.kt:302-328, inflate/InflateStream.kt:376-438, bitwise/checksum/Adler32Utils.kt:31-57),
// every byte of work done behind the C ABI of include/zlibk_gpu.h.  No JNI-style fallback, no CPU path: if the
// library or a device is missing, construction fails.
@file:OptIn(kotlinx.cinterop.ExperimentalForeignApi::class)

package ai.solace.zlib.gpu

import ai.solace.zlib.gpu.cinterop.*
import kotlinx.cinterop.*
import okio.Buffer
import okio.BufferedSink
import okio.BufferedSource

class ZStreamException(message: String) : Exception(message)

/** One engine per thread ("not thread-safe, one ZStream per thread", docs/API.md:725). */
class GpuEngine(device: Int = 0) {
    internal val ctx: CPointer<zk_ctx> =
        zk_ctx_create(device) ?: throw ZStreamException("no B200 engine: ${zk_last_error()?.toKString()}")

    fun close() = zk_ctx_destroy(ctx)
}

private val engine by lazy { GpuEngine(0) }

/**
 * Page-locked slabs for the host-buffer calls.  A Kotlin ByteArray.usePinned is pinned for the garbage collector only: to
 * CUDA it is pageable memory, every copy is staged through the driver and the engine's copy / compute overlap is lost
 * (8 GB/s instead of 41 GB/s, profiles/pageable_r2.md).  The adapters therefore move the data of an Okio stream through
 * slabs from zk_host_alloc, at most SLAB bytes at a time: no 2 GiB ByteArray ceiling, and the sizes of BASELINE.json
 * configs[2] and [3] (4 GiB, 8 GiB) pass through in 256 MiB steps.
 */
private const val SLAB = 256L shl 20
private const val WINDOW = 32768

private class PinnedSlab(val bytes: Long) {
    val ptr: CPointer<UByteVar> = zk_host_alloc(bytes.toULong())?.reinterpret()
        ?: throw ZStreamException("zk_host_alloc($bytes) failed")
    fun free() = zk_host_free(ptr)
}

/** Reads up to `max` bytes of the source into pinned memory at `dst`; returns the count (0 at the end of the source). */
private fun BufferedSource.readInto(dst: CPointer<UByteVar>, max: Long): Long {
    val chunk = ByteArray(1 shl 20)
    var got = 0L
    while (got < max) {
        val n = read(chunk, 0, minOf(chunk.size.toLong(), max - got).toInt())
        if (n <= 0) break
        chunk.usePinned { platform.posix.memcpy(dst + got, it.addressOf(0), n.toULong()) }
        got += n
    }
    return got
}

private fun BufferedSink.writeFrom(src: CPointer<UByteVar>, count: Long) {
    val chunk = ByteArray(1 shl 20)
    var at = 0L
    while (at < count) {
        val n = minOf(chunk.size.toLong(), count - at).toInt()
        chunk.usePinned { platform.posix.memcpy(it.addressOf(0), src + at, n.toULong()) }
        write(chunk, 0, n)
        at += n
    }
}

object DeflateStream {
    /** DeflateStream.compressZlib (DS:302-311): returns the bytes consumed. */
    fun compressZlib(source: BufferedSource, sink: BufferedSink, level: Int = 6): Long {
        val (rc, n) = compressZlibResult(source, sink, level)
        if (rc != ZK_OK) throw ZStreamException("compressZlib failed: $rc")
        return n
    }

    /**
     * DeflateStream.compressZlibResult (DS:318-328): (Z_STREAM_ERROR, 0) and an untouched sink for level > 9.
     * One zlib stream made of slab bodies: every slab is compressed as a headerless, byte-aligned body primed with the
     * 32 KiB before it (zk_deflate_shard, the multi-GPU sharding primitive), the checksums fold with zk_adler32_combine,
     * the header goes in front and the final block + Adler-32 behind (zk_zlib_header / zk_zlib_tail).
     */
    fun compressZlibResult(source: BufferedSource, sink: BufferedSink, level: Int = 6): Pair<Int, Long> {
        if (level > 9) return ZK_STREAM_ERROR to 0L
        val input = PinnedSlab(WINDOW + SLAB)          // [32 KiB of history | slab]
        val output = PinnedSlab(zk_compress_bound(SLAB.toULong()).toLong())
        try {
            memScoped {
                val head = allocArray<UByteVar>(2)
                zk_zlib_header(level, head)
                sink.writeFrom(head, 2)
                val bodyLen = alloc<ULongVar>()
                val slabAdler = alloc<UIntVar>()
                var adler = 1u
                var total = 0L
                var hist = 0
                while (true) {
                    val n = source.readInto((input.ptr + WINDOW)!!, SLAB)
                    if (n == 0L) break
                    val rc = zk_deflate_shard(
                        engine.ctx, if (hist > 0) (input.ptr + (WINDOW - hist)) else null, hist.toUInt(),
                        input.ptr + WINDOW, n.toULong(), level, ZK_DEFAULT_STRATEGY,
                        output.ptr, output.bytes.toULong(), bodyLen.ptr, slabAdler.ptr,
                    )
                    if (rc != ZK_OK) return rc to total
                    sink.writeFrom(output.ptr, bodyLen.value.toLong())
                    adler = zk_adler32_combine(adler, slabAdler.value, n.toULong())
                    total += n
                    // the last 32 KiB of this slab prime the next one
                    hist = minOf(WINDOW.toLong(), if (n >= WINDOW) WINDOW.toLong() else hist + n).toInt()
                    platform.posix.memmove(input.ptr + (WINDOW - hist), input.ptr + (WINDOW + n - hist), hist.toULong())
                    if (n < SLAB) break
                }
                val tail = allocArray<UByteVar>(6)
                zk_zlib_tail(adler, tail)
                sink.writeFrom(tail, 6)
                sink.flush()
                return ZK_OK to total
            }
        } finally {
            input.free()
            output.free()
        }
    }
}

object InflateStream {
    /**
     * InflateStream.inflateZlib (IS:376-438): (Z_STREAM_END, bytesOut) on success, the bytes written so far beside any
     * other code.  The stream goes through the inflate handle slab by slab (zk_inflate carries the 32 KiB window, the
     * unconsumed tail, the checksum and the totals from call to call and decodes every block once).
     */
    fun inflateZlib(source: BufferedSource, sink: BufferedSink): Pair<Int, Long> {
        val h = zk_inflate_init(engine.ctx) ?: return ZK_MEM_ERROR to 0L
        val input = PinnedSlab(SLAB / 4)
        val output = PinnedSlab(SLAB)
        var written = 0L
        try {
            memScoped {
                val inLen = alloc<ULongVar>()
                val outLen = alloc<ULongVar>()
                var ended = false
                while (true) {
                    val n = if (ended) 0L else source.readInto(input.ptr, input.bytes)
                    if (n < input.bytes) ended = true
                    var fed = 0L
                    do {  // hand the slab over, then drain what the engine has decoded
                        inLen.value = (n - fed).toULong()
                        outLen.value = output.bytes.toULong()
                        val rc = zk_inflate(h, if (n > fed) input.ptr + fed else null, inLen.ptr, output.ptr, outLen.ptr,
                                            if (ended) ZK_FINISH else ZK_NO_FLUSH)
                        fed += inLen.value.toLong()
                        if (outLen.value > 0uL) {
                            sink.writeFrom(output.ptr, outLen.value.toLong())
                            written += outLen.value.toLong()
                        }
                        if (rc == ZK_STREAM_END) { sink.flush(); return ZK_STREAM_END to written }
                        if (rc != ZK_OK && !(rc == ZK_BUF_ERROR && !ended)) { sink.flush(); return rc to written }  // IS:425-437
                    } while (fed < n || outLen.value == output.bytes.toULong())
                    if (ended && n == 0L) { sink.flush(); return ZK_BUF_ERROR to written }
                }
            }
        } finally {
            zk_stream_end(h)
            input.free()
            output.free()
        }
        @Suppress("UNREACHABLE_CODE")
        return ZK_STREAM_ERROR to written
    }
}

object Adler32Utils {
    /** Adler32Utils.adler32 (AU:31-57): a null buffer resets to 1. */
    fun adler32(adler: Long, buf: ByteArray?, index: Int, len: Int): Long {
        if (buf == null) return 1L
        if (len == 0) return adler and 0xFFFFFFFFL
        return buf.usePinned { p ->
            zk_adler32(engine.ctx, adler.toUInt(), p.addressOf(index).reinterpret(), len.toULong()).toLong()
        }
    }
}

/** ZLib.compress / ZLib.decompress (README.md:84-92). */
object ZLib {
    fun compress(data: ByteArray, level: Int = 6): ByteArray {
        val sink = Buffer()
        val (rc, _) = DeflateStream.compressZlibResult(Buffer().write(data), sink, level)
        if (rc != ZK_OK) throw ZStreamException("compress failed: $rc")
        return sink.readByteArray()
    }

    fun decompress(data: ByteArray): ByteArray {
        val sink = Buffer()
        val (rc, _) = InflateStream.inflateZlib(Buffer().write(data), sink)
        if (rc != ZK_STREAM_END) throw ZStreamException("decompress failed: $rc")
        return sink.readByteArray()
    }
}

/** The documented ZStream (docs/API.md:205-281) over the streaming handles of the C ABI. */
class ZStream {
    var next_in: ByteArray = ByteArray(0)
    var next_in_index = 0
    var avail_in = 0
    var total_in = 0L
    var next_out: ByteArray = ByteArray(0)
    var next_out_index = 0
    var avail_out = 0
    var total_out = 0L
    var msg: String? = null
    var adler = 1L
    var data_type = 2 // Z_UNKNOWN

    private var handle: CPointer<zk_stream>? = null
    private var deflating = false

    fun deflateInit(level: Int = 6, method: Int = 8, windowBits: Int = 15, memLevel: Int = 8, strategy: Int = 0): Int {
        if (method != 8 || level > 9) return ZK_STREAM_ERROR
        end()
        val container = when (windowBits) { in 8..15 -> 0; in -15..-8 -> 1; in 24..31 -> 2; else -> return ZK_STREAM_ERROR }
        handle = zk_deflate_init2(engine.ctx, level, strategy, container)
        deflating = true
        return if (handle != null) ZK_OK else ZK_STREAM_ERROR
    }

    fun inflateInit(windowBits: Int = 15): Int {
        end()
        val container = when (windowBits) { in 8..15 -> 0; in -15..-8 -> 1; in 24..31 -> 2; else -> return ZK_STREAM_ERROR }
        handle = zk_inflate_init2(engine.ctx, container)
        deflating = false
        return if (handle != null) ZK_OK else ZK_STREAM_ERROR
    }

    fun deflate(flush: Int = ZK_NO_FLUSH): Int = run(flush, true)
    fun inflate(flush: Int = ZK_NO_FLUSH): Int = run(flush, false)
    fun deflateParams(level: Int, strategy: Int): Int = handle?.let { zk_deflate_params(it, level, strategy) } ?: ZK_STREAM_ERROR
    fun deflateSetDictionary(dictionary: ByteArray, dictLength: Int = dictionary.size): Int =
        handle?.let { h -> dictionary.usePinned { zk_deflate_set_dictionary(h, it.addressOf(0).reinterpret(), dictLength.toUInt()) } }
            ?: ZK_STREAM_ERROR
    fun inflateSetDictionary(dictionary: ByteArray, dictLength: Int = dictionary.size): Int =
        handle?.let { h -> dictionary.usePinned { zk_inflate_set_dictionary(h, it.addressOf(0).reinterpret(), dictLength.toUInt()) } }
            ?: ZK_STREAM_ERROR
    fun deflateCopy(dest: ZStream): Int {
        val h = handle ?: return ZK_STREAM_ERROR
        if (!deflating) return ZK_STREAM_ERROR
        dest.end()
        dest.handle = zk_deflate_copy(h) ?: return ZK_MEM_ERROR
        dest.deflating = true
        dest.next_in = next_in; dest.next_in_index = next_in_index; dest.avail_in = avail_in; dest.total_in = total_in
        dest.next_out = next_out.copyOf(); dest.next_out_index = next_out_index; dest.avail_out = avail_out; dest.total_out = total_out
        dest.adler = adler
        return ZK_OK
    }
    fun deflateReset(): Int = handle?.let { zk_deflate_reset(it) } ?: ZK_STREAM_ERROR
    fun inflateReset(): Int = handle?.let { zk_inflate_reset(it) } ?: ZK_STREAM_ERROR
    fun inflateSync(): Int = ZK_DATA_ERROR
    fun deflateEnd(): Int = end()
    fun inflateEnd(): Int = end()
    fun free() { end(); next_in = ByteArray(0); next_out = ByteArray(0) }

    private fun run(flush: Int, wantDeflate: Boolean): Int {
        val h = handle ?: return ZK_STREAM_ERROR
        if (wantDeflate != deflating) return ZK_STREAM_ERROR
        memScoped {
            val inLen = alloc<ULongVar>().apply { value = avail_in.toULong() }
            val outLen = alloc<ULongVar>().apply { value = avail_out.toULong() }
            val rc = next_in.usePinned { i ->
                next_out.usePinned { o ->
                    val ip = if (avail_in > 0) i.addressOf(next_in_index).reinterpret<UByteVar>() else null
                    val op = if (avail_out > 0) o.addressOf(next_out_index).reinterpret<UByteVar>() else null
                    if (wantDeflate) zk_deflate(h, ip, inLen.ptr, op, outLen.ptr, flush)
                    else zk_inflate(h, ip, inLen.ptr, op, outLen.ptr, flush)
                }
            }
            next_in_index += inLen.value.toInt(); avail_in -= inLen.value.toInt()
            next_out_index += outLen.value.toInt(); avail_out -= outLen.value.toInt()
            total_in = zk_stream_total_in(h).toLong()
            total_out = zk_stream_total_out(h).toLong()
            adler = zk_stream_adler(h).toLong()
            if (rc < 0) msg = "zlibk error $rc"
            return rc
        }
    }

    private fun end(): Int {
        val h = handle ?: return ZK_STREAM_ERROR
        zk_stream_end(h)
        handle = null
        return ZK_OK
    }
}

/**
 * gzip members and CRC-32 (SURVEY.md 8f row 4).  The reference has neither (only a logging stub,
 * common/ZlibLogger.kt:115-118); the shapes follow zlib.h / RFC 1952 so that `pigz` output (gzip by default) can be
 * read and written next to the zlib path.
 */
object Gzip {
    fun compress(data: ByteArray, level: Int = 6): ByteArray {
        val out = ByteArray(zk_compress_bound(data.size.toULong()).toInt() + 32)
        memScoped {
            val n = alloc<ULongVar>()
            val rc = data.usePinned { i ->
                out.usePinned { o ->
                    zk_deflate_gzip(engine.ctx, if (data.isEmpty()) null else i.addressOf(0).reinterpret(), data.size.toULong(), level,
                                    ZK_DEFAULT_STRATEGY, o.addressOf(0).reinterpret(), out.size.toULong(), n.ptr, null)
                }
            }
            if (rc != ZK_OK) throw ZStreamException("gzip compress failed: $rc")
            return out.copyOf(n.value.toInt())
        }
    }

    /** All members of the file, outputs concatenated (RFC 1952 2.2). */
    fun decompress(data: ByteArray, sizeHint: Int = data.size * 8 + 65536): ByteArray {
        var cap = sizeHint
        while (true) {
            val out = ByteArray(cap)
            memScoped {
                val n = alloc<ULongVar>()
                val used = alloc<ULongVar>()
                val rc = data.usePinned { i ->
                    out.usePinned { o ->
                        zk_inflate_gzip(engine.ctx, i.addressOf(0).reinterpret(), data.size.toULong(), o.addressOf(0).reinterpret(),
                                        cap.toULong(), n.ptr, used.ptr)
                    }
                }
                if (rc == ZK_STREAM_END) return out.copyOf(n.value.toInt())
                if (rc != ZK_BUF_ERROR || cap >= Int.MAX_VALUE / 4) throw ZStreamException("gzip decompress failed: $rc")
            }
            cap *= 4
        }
    }
}

object Crc32 {
    /** zlib.h crc32(crc, buf, len); a null buffer gives the initial value 0. */
    fun crc32(crc: Long, buf: ByteArray?, index: Int = 0, len: Int = buf?.size ?: 0): Long {
        if (buf == null) return 0L
        if (len == 0) return crc
        return buf.usePinned { p -> zk_crc32(engine.ctx, crc.toUInt(), p.addressOf(index).reinterpret(), len.toULong()).toLong() }
    }

    fun combine(crc1: Long, crc2: Long, len2: Long): Long = zk_crc32_combine(crc1.toUInt(), crc2.toUInt(), len2.toULong()).toLong()
}
