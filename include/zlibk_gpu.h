/*
 * zlibk_gpu.h -- C ABI of the B200 (sm_100a) DEFLATE engine that drops in behind
 * ZLib.kotlin's compression hot path.
 *
 * Every entry point is extern "C", takes plain pointers and 64-bit sizes, returns a
 * zlib code (common/Constants.kt:98-106), never throws and keeps no global state
 * beyond the per-thread default context.  One zk_ctx per host thread ("not
 * thread-safe, one ZStream per thread", docs/API.md:725); contexts are independent.
 * There is NO CPU fallback: without a CUDA device every compute entry point returns
 * ZK_ERRNO / a null context.
 *
 * Citations are into /root/reference/src/commonMain/kotlin/ai/solace/zlib/ unless a
 * path is given.  The Kotlin/Native cinterop binding a maintainer would add is in
 * INTEGRATION.md (zlib.kotlin_b200/kotlin/zlibk.def + ZlibkGpu.kt).
 */
#ifndef ZLIBK_GPU_H
#define ZLIBK_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- codes and constants (common/Constants.kt) ----------------------------- */
#define ZK_OK 0             /* K:98  */
#define ZK_STREAM_END 1     /* K:99  */
#define ZK_NEED_DICT 2      /* K:100 */
#define ZK_ERRNO (-1)       /* K:101: CUDA runtime failure, no device */
#define ZK_STREAM_ERROR (-2)/* K:102: bad parameter / bad handle state */
#define ZK_DATA_ERROR (-3)  /* K:103 */
#define ZK_MEM_ERROR (-4)   /* K:104: device or pinned allocation failed */
#define ZK_BUF_ERROR (-5)   /* K:105: output too small / input truncated */
#define ZK_VERSION_ERROR (-6)

#define ZK_NO_COMPRESSION 0        /* K:54 */
#define ZK_BEST_SPEED 1            /* K:55 */
#define ZK_BEST_COMPRESSION 9      /* K:56 */
#define ZK_DEFAULT_COMPRESSION (-1)/* K:57 -- the reference maps every level <= 0 to stored (DS:307-311) */

#define ZK_DEFAULT_STRATEGY 0 /* K:66 */
#define ZK_FILTERED 1         /* K:67 */
#define ZK_HUFFMAN_ONLY 2     /* K:68 */

#define ZK_NO_FLUSH 0      /* K:79 */
#define ZK_PARTIAL_FLUSH 1 /* K:80 */
#define ZK_SYNC_FLUSH 2    /* K:81 */
#define ZK_FULL_FLUSH 3    /* K:82 */
#define ZK_FINISH 4        /* K:83 */

typedef struct zk_ctx zk_ctx;       /* device, stream, scratch arena */
typedef struct zk_stream zk_stream; /* streaming deflate / inflate state */

/* ---- library ---------------------------------------------------------------- */
/* replaces version(): String, common/Constants.kt:44 ("1.0.2"), suffixed with the engine tag */
const char *zk_version(void);
/* number of CUDA devices visible (0 = none: the engine is unusable) */
int zk_device_count(void);
/* text of the last CUDA error seen by this thread ("" if none) */
const char *zk_last_error(void);

/* ---- contexts ----------------------------------------------------------------- */
/* Create a context bound to CUDA device `device` with its own stream.  Returns NULL
 * (and sets zk_last_error) when the device cannot be used. */
zk_ctx *zk_ctx_create(int device);
void zk_ctx_destroy(zk_ctx *ctx);
/* Run the context's kernels on a caller-owned cudaStream_t (e.g. torch's current
 * stream) instead of its own; pass NULL to restore the private stream. */
int zk_ctx_set_stream(zk_ctx *ctx, void *cuda_stream);
int zk_ctx_synchronize(zk_ctx *ctx);
/* Tunables (0 restores the built-in value unless noted): "chunk_bytes" (LZ unit, default auto),
 * "block_bytes" (deflate block span, a power of two in [4096, 32768], default 32768),
 * "slab_bytes" (slab of a device-resident deflate; default: two equal slabs from 512 MiB to 2 GiB, 1 GiB beyond),
 * "host_slab_bytes" / "host_first_slab_bytes" (slabs of the pipelined host-buffer deflate, default 64 MiB / 32 MiB;
 * multiples of 64 KiB, at least 1 MiB), "slab_overlap" (default 1: slabs alternate between two streams; 0: in series),
 * "lz_chain" / "lz_max_lazy" / "lz_nice" (override the level table, -1 = off),
 * "auto_register" (1: a host-buffer call page-locks caller buffers of 1 MiB or more that CUDA does not know yet, once, and
 * keeps them locked until the context is destroyed -- for callers that keep and reuse their buffers, up to 16 of them; the
 * caller must not free such a buffer while the context lives), "inflate_pipe" / "inflate_slab_bytes" (opt-in slab pipeline
 * of the host-buffer inflate), "inflate_foreign" (default 1: a stream without sync markers -- libz, gzip tools, the reference --
 * is cut at the dynamic block headers the engine finds in it and decoded in parallel; 0: one warp decodes it),
 * "inflate_prefix_bytes" (stream handle: gathered bytes of a stream without markers from which everything up to the last
 * entry point is decoded and handed out; 0 = 32 MiB, at least 256 KiB),
 * "timing" (1: CUDA events around every launch, read with the "*_us" counters; a profiling aid that costs the slab
 * overlap about a tenth of its throughput). */
int zk_ctx_set_option(zk_ctx *ctx, const char *name, int64_t value);
/* Counters since creation: name in {"kernel_launches","h2d_bytes","d2h_bytes","inflate_foreign_chunks",
 * "inflate_foreign_merges","inflate_prefix_decodes","inflate_quick_incomplete","inflate_pipe_slabs","inflate_pipe_fallbacks"}; with "timing" on also
 * {"lz77","huffman","pack","adler","inflate_decode","inflate_other"} + "_us" / "_spans"; "reset_timing" clears those. */
int64_t zk_ctx_get_counter(zk_ctx *ctx, const char *name);

/* ---- caller buffers the copy engines can stream from ------------------------------------
 * The host-buffer calls below overlap their copies with the kernels only when the caller's memory is page-locked for
 * CUDA.  A Kotlin ByteArray.usePinned / JVM critical array is pinned for the garbage collector only: to CUDA it is
 * pageable, every copy is staged through the driver and the overlap is lost (measured: profiles/pageable_r2.md).
 * zk_host_alloc hands out page-locked memory to keep and reuse (nativeHeap-style, free with zk_host_free);
 * zk_host_register page-locks a buffer the caller already owns for as long as it lives (costly: do it once per
 * buffer, not once per call), zk_host_unregister undoes it.  All return ZK_OK / ZK_MEM_ERROR / ZK_STREAM_ERROR. */
void *zk_host_alloc(uint64_t bytes);
void zk_host_free(void *p);
int zk_host_register(void *p, uint64_t bytes);
int zk_host_unregister(void *p);

/* ---- Adler-32 ------------------------------------------------------------------ */
/* replaces Adler32Utils.adler32(adler, buf, index, len), bitwise/checksum/Adler32Utils.kt:31-57
 * (and the wrapper deflate/Adler32.kt:35).  buf is HOST memory; buf == NULL -> 1 (AU:37). */
uint32_t zk_adler32(zk_ctx *ctx, uint32_t adler, const uint8_t *buf, uint64_t len);
/* same on a DEVICE buffer, asynchronous result written to *d_out (device uint32_t[1]) */
int zk_adler32_device(zk_ctx *ctx, uint32_t adler, const uint8_t *d_buf, uint64_t len, uint32_t *d_out);
/* zlib's adler32_combine: checksum of A||B from adler(A), adler(B), len(B).  Pure integer
 * arithmetic on three scalars, used to fold per-shard checksums. */
uint32_t zk_adler32_combine(uint32_t adler1, uint32_t adler2, uint64_t len2);

/* ---- one-shot zlib deflate / inflate, HOST buffers ------------------------------ */
/* upper bound of the compressed size for n input bytes at any level */
uint64_t zk_compress_bound(uint64_t n);

/* replaces DeflateStream.compressZlibResult(source, sink, level), deflate/DeflateStream.kt:302-328.
 *   level > 9            -> ZK_STREAM_ERROR, *out_len = 0, nothing written (DS:323-325)
 *   level <= 0           -> stored blocks (DS:307-308)
 *   level 1 / 2..9       -> fastest / dynamic-Huffman blocks; FLEVEL per DS:27-33
 *   strategy             -> zlib.h semantics (the reference defines the constants only, K:66-68)
 * Returns ZK_OK, ZK_BUF_ERROR if cap < needed (nothing usable in out), or an error code.
 * *adler (optional) receives the Adler-32 of the input. */
int zk_deflate_zlib(zk_ctx *ctx, const uint8_t *in, uint64_t n, int level, int strategy, uint8_t *out,
                    uint64_t cap, uint64_t *out_len, uint32_t *adler);

/* replaces InflateStream.inflateZlib(source, sink), inflate/InflateStream.kt:376-438.
 * Returns ZK_STREAM_END on success (IS:424 -- not ZK_OK), ZK_NEED_DICT for FDICT streams
 * (IS:36-43, *consumed = 6), ZK_DATA_ERROR, ZK_BUF_ERROR (truncated input, IS:425-428, or
 * cap too small).  *out_len = bytes written, *consumed = input bytes used (trailing bytes
 * of a second member are left unread, as in the reference).  The two exceptions the
 * reference lets escape (IS:243 -> CH:127,133) are reported as ZK_DATA_ERROR. */
int zk_inflate_zlib(zk_ctx *ctx, const uint8_t *in, uint64_t m, uint8_t *out, uint64_t cap, uint64_t *out_len,
                    uint64_t *consumed);

/* ---- one-shot, DEVICE buffers (no host copies; asynchronous on the context stream) -- */
/* d_out must hold zk_compress_bound(n) bytes.  d_result (device uint64_t[4]) receives
 * {out_len, adler32, status, n}.  Call zk_ctx_synchronize before reading it. */
int zk_deflate_zlib_device(zk_ctx *ctx, const uint8_t *d_in, uint64_t n, int level, int strategy, uint8_t *d_out,
                           uint64_t cap, uint64_t *d_result);
/* d_result (device uint64_t[4]) receives {out_len, consumed, status, adler32}. */
int zk_inflate_zlib_device(zk_ctx *ctx, const uint8_t *d_in, uint64_t m, uint8_t *d_out, uint64_t cap,
                           uint64_t *d_result);

/* ---- preset dictionaries (zlib.h deflateSetDictionary / inflateSetDictionary) ----
 * The reference only detects FDICT and returns Z_NEED_DICT (inflate/InflateStream.kt:36-43); these entry points take
 * the stream further: deflate emits FDICT + DICTID (Adler-32 of the dictionary) and primes matching with the
 * dictionary's last 32 KiB; inflate checks DICTID (ZK_DATA_ERROR on mismatch) and resolves references into it. */
int zk_deflate_zlib_dict(zk_ctx *ctx, const uint8_t *dict, uint32_t dict_len, const uint8_t *in, uint64_t n, int level,
                         int strategy, uint8_t *out, uint64_t cap, uint64_t *out_len, uint32_t *adler);
int zk_inflate_zlib_dict(zk_ctx *ctx, const uint8_t *in, uint64_t m, const uint8_t *dict, uint32_t dict_len, uint8_t *out,
                         uint64_t cap, uint64_t *out_len, uint64_t *consumed);

/* ---- CRC-32 and the gzip / raw-deflate containers ----------------
 * The reference has neither (only a logging stub, common/ZlibLogger.kt:115-118); pigz's default format is gzip.
 * zlib.h semantics: crc32(crc, buf, len), crc32_combine; RFC 1952 members, RFC 1951 raw streams. */
uint32_t zk_crc32(zk_ctx *ctx, uint32_t crc, const uint8_t *buf, uint64_t len);
int zk_crc32_device(zk_ctx *ctx, uint32_t crc, const uint8_t *d_buf, uint64_t len, uint32_t *d_out);
uint32_t zk_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);
int zk_deflate_gzip(zk_ctx *ctx, const uint8_t *in, uint64_t n, int level, int strategy, uint8_t *out, uint64_t cap,
                    uint64_t *out_len, uint32_t *crc);
/* every member of the file is decoded, outputs concatenated (RFC 1952 2.2) */
int zk_inflate_gzip(zk_ctx *ctx, const uint8_t *in, uint64_t m, uint8_t *out, uint64_t cap, uint64_t *out_len,
                    uint64_t *consumed);
int zk_deflate_raw(zk_ctx *ctx, const uint8_t *in, uint64_t n, int level, int strategy, uint8_t *out, uint64_t cap,
                   uint64_t *out_len);
int zk_inflate_raw(zk_ctx *ctx, const uint8_t *in, uint64_t m, uint8_t *out, uint64_t cap, uint64_t *out_len,
                   uint64_t *consumed);

/* ---- sharded deflate across GPUs ------------------------------------
 * Each rank compresses a contiguous range of the input as a headerless, byte-aligned body
 * (every block ends with a sync marker).  Ranks exchange {body_len, adler, raw_len} with
 * ONE allgather (NCCL, done by the caller), then every rank knows its global offset and
 * the folded checksum.  Rank 0 prepends the 2-byte header, the last rank appends the final
 * empty block and the trailer. */
/* d_hist: the 32 KiB of input that precede this shard (device), hist_len in [0, 32768].
 * d_meta (device uint64_t[4]) receives {body_len, adler32(shard), status, n} -- the same record layout as
 * zk_deflate_zlib_device's d_result. */
int zk_deflate_shard_device(zk_ctx *ctx, const uint8_t *d_hist, uint32_t hist_len, const uint8_t *d_in, uint64_t n,
                            int level, int strategy, uint8_t *d_body, uint64_t cap, uint64_t *d_meta);
/* The same with host buffers (pipelined H2D / compute / D2H like zk_deflate_zlib): `hist` = the up to 32 KiB of input
 * before this shard (host), the body goes to out_body. */
int zk_deflate_shard(zk_ctx *ctx, const uint8_t *hist, uint32_t hist_len, const uint8_t *in, uint64_t n, int level,
                     int strategy, uint8_t *out_body, uint64_t cap, uint64_t *body_len, uint32_t *adler);
/* One stream inflated across GPUs: every rank decodes the body of its chunk range (phase 1), the ranks all_gather their
 * 64 KiB tail maps (the only collective), each composes the maps of the ranks before it into the 32 KiB in front of its
 * shard, and phase 2 resolves the shard and writes its bytes.  hist_len = how far back the shard may refer (0 for the
 * first rank, 32768 otherwise).  d_result (uint64_t[4]) = {out_len, consumed, status, Adler-32 of the shard's output};
 * phase 1 reports ZK_OK when the body decoded, phase 2 ZK_STREAM_END.  Phase 2 uses scratch phase 1 left in the
 * context: no other call on the context in between. */
int zk_inflate_shard_prepare_device(zk_ctx *ctx, const uint8_t *d_body, uint64_t m, uint32_t hist_len, uint64_t cap,
                                    uint16_t *d_tail_map, uint64_t *d_result);
int zk_window_from_maps_device(zk_ctx *ctx, const uint16_t *d_maps, uint64_t stride_bytes, uint32_t count, uint8_t *d_window);
int zk_inflate_shard_finish_device(zk_ctx *ctx, const uint8_t *d_entry_window, uint8_t *d_out, uint64_t *d_result);
/* zlib header for `level` (2 bytes, DS:35-50) and stream tail (final empty block + big-endian
 * Adler-32, 6 bytes) as host helpers for the rank that owns them. */
void zk_zlib_header(int level, uint8_t out2[2]);
void zk_zlib_tail(uint32_t adler, uint8_t out6[6]);

/* ---- streaming (the documented ZStream surface, docs/API.md:205-281) ---------------- */
/* deflateInit(level, method=8, windowBits=15, memLevel, strategy) */
zk_stream *zk_deflate_init(zk_ctx *ctx, int level, int strategy);
/* deflate(flush): consumes *in_len bytes from `in`, writes at most *out_len bytes to `out`;
 * both are updated to the amounts used.  zlib.h flush semantics. */
int zk_deflate(zk_stream *s, const uint8_t *in, uint64_t *in_len, uint8_t *out, uint64_t *out_len, int flush);
int zk_deflate_reset(zk_stream *s);
int zk_deflate_params(zk_stream *s, int level, int strategy);
int zk_deflate_set_dictionary(zk_stream *s, const uint8_t *dict, uint32_t len);
/* inflateInit(windowBits=15) */
zk_stream *zk_inflate_init(zk_ctx *ctx);
/* zlib.h deflateInit2 / inflateInit2 windowBits variants: container 0 = zlib (windowBits 8..15), 1 = raw deflate
 * (windowBits < 0), 2 = gzip (windowBits + 16).  zk_deflate_copy = deflateCopy (docs/API.md:253). */
zk_stream *zk_deflate_init2(zk_ctx *ctx, int level, int strategy, int container);
zk_stream *zk_inflate_init2(zk_ctx *ctx, int container);
zk_stream *zk_deflate_copy(const zk_stream *src);
int zk_inflate(zk_stream *s, const uint8_t *in, uint64_t *in_len, uint8_t *out, uint64_t *out_len, int flush);
int zk_inflate_reset(zk_stream *s);
int zk_inflate_set_dictionary(zk_stream *s, const uint8_t *dict, uint32_t len);
/* totals and running checksum (ZStream.total_in / total_out / adler, docs/API.md:217-234) */
uint64_t zk_stream_total_in(const zk_stream *s);
uint64_t zk_stream_total_out(const zk_stream *s);
uint32_t zk_stream_adler(const zk_stream *s);
/* deflateEnd / inflateEnd / free */
void zk_stream_end(zk_stream *s);

#ifdef __cplusplus
}
#endif
#endif
