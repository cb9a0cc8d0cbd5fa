/*
 * zk_oracle.h -- CPU restatement of the ZLib.kotlin reference algorithms.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * the reported CPU baseline.  The product path (zlib.kotlin_b200/csrc) never
 * links or calls it.
 *
 * Parity pinning: the reference (Kotlin/Native) cannot be built in this
 * image, so the restatement is pinned against
 *   - every Adler-32 known-answer value in the reference's tests,
 *   - the four inflate error streams + the "hello world" round trip,
 *   - the deflate parameter tests and the 8-byte empty-input stream,
 * and cross-checked against libz 1.3 (tests/test_oracle_*.py).  Deflate
 * OUTPUT BYTES of the reference are not pinned by any reference test
 * ("parity unpinned" for compressed bytes / ratio).
 *
 * Citations are relative to /root/reference/src/commonMain/kotlin/ai/solace/zlib/.
 */
#ifndef ZK_ORACLE_H
#define ZK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* zlib return codes, common/Constants.kt:98-106 */
#define ZKO_OK 0
#define ZKO_STREAM_END 1
#define ZKO_NEED_DICT 2
#define ZKO_ERRNO (-1)
#define ZKO_STREAM_ERROR (-2)
#define ZKO_DATA_ERROR (-3)
#define ZKO_MEM_ERROR (-4)
#define ZKO_BUF_ERROR (-5)
/* Not a zlib code: the reference lets an IllegalStateException escape
 * inflateZlib (CanonicalHuffman.kt:127,133 reached from InflateStream.kt:243,315);
 * the oracle reports that case with this value. */
#define ZKO_ESCAPED_EXCEPTION (-100)

/* bitwise/checksum/Adler32Utils.kt:31-57.  buf == NULL -> 1 (AU:37). */
uint32_t zko_adler32(uint32_t adler, const uint8_t *buf, size_t len);

/*
 * inflate/InflateStream.kt:376-438 (inflateZlib).
 *   out/out_cap : sink; bytes the reference would have written to its sink
 *                 (including bytes written before an error is detected)
 *   *out_len    : number of bytes written to the sink
 *   *reported   : the Long the reference returns as Pair.second
 *   *consumed   : input bytes pulled from the source (the bit reader is
 *                 byte-granular, BR:48-57)
 * Returns the Int the reference returns as Pair.first, ZKO_ESCAPED_EXCEPTION
 * where it would throw, or ZKO_MEM_ERROR if out_cap is too small (oracle-only
 * condition; the reference sink is unbounded).
 */
int zko_inflate_zlib(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                     size_t *out_len, uint64_t *reported, size_t *consumed);

/*
 * deflate/DeflateStream.kt:302-328 (compressZlib / compressZlibResult) with
 * every behavioural quirk of the reference kept,
 * including the Okio 8 KiB read granularity.  Output buffer is malloc'd;
 * release it with zko_free.  *bytes_in is the Long compressZlib returns.
 * Returns ZKO_OK, or ZKO_STREAM_ERROR (nothing written) for level > 9.
 */
int zko_deflate_zlib(const uint8_t *in, size_t n, int level, uint8_t **out, size_t *out_len,
                     uint64_t *bytes_in);
void zko_free(void *p);

/* deflate/HuffmanBuilder.kt:10-83 (buildLengths).  ensure_symbol < 0 == null. */
void zko_build_lengths(const int *freq, int n, int max_bits, int ensure_symbol, int *out_lengths);

/* inflate/CanonicalHuffman.kt:103-120 (buildEncoder): codes are bit-reversed (LSB-first). */
void zko_build_encoder(const int *lengths, int n, int *codes, int *bits);

/*
 * inflate/CanonicalHuffman.kt:54-82 (buildFull).  Fills bits_tab/vals_tab with
 * 2^maxLen entries (caller provides 32768 each).  Returns maxLen, or -1 when
 * the reference's FullTable constructor would throw (maxLen > 15).
 */
int zko_build_full(const int *lengths, int n, int *bits_tab, int *vals_tab);

/* common/Constants.kt:37-44 */
const char *zko_version(void);

#ifdef __cplusplus
}
#endif
#endif
