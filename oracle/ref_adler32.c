/*
 * ref_adler32.c -- oracle restatement of Adler32Utils.adler32.
 * TEST INFRASTRUCTURE ONLY (see zk_oracle.h).
 *
 * Follows bitwise/checksum/Adler32Utils.kt:31-57 with ADLER_BASE / ADLER_NMAX
 * from common/Constants.kt:110-111: the two running sums are reduced modulo
 * 65521 once per run of at most 5552 bytes.
 */
#include "zk_oracle.h"

enum { kBase = 65521, kNmax = 5552 };

uint32_t zko_adler32(uint32_t adler, const uint8_t *buf, size_t len) {
    if (buf == NULL) return 1u; /* AU:37 -- a null buffer resets to 1 whatever the seed */
    uint64_t lo = adler & 0xFFFFu;         /* AU:39 */
    uint64_t hi = (adler >> 16) & 0xFFFFu; /* AU:40 */
    size_t done = 0;
    while (done < len) { /* AU:45-54 */
        size_t run = len - done;
        if (run > kNmax) run = kNmax;
        for (size_t k = 0; k < run; ++k) {
            lo += buf[done + k];
            hi += lo;
        }
        done += run;
        lo %= kBase;
        hi %= kBase;
    }
    return (uint32_t)(((hi & 0xFFFFu) << 16) | (lo & 0xFFFFu)); /* AU:56 */
}

const char *zko_version(void) { return "1.0.2"; /* common/Constants.kt:37-44 */ }
