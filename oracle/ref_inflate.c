/*
 * ref_inflate.c -- oracle restatement of InflateStream.inflateZlib.
 * TEST INFRASTRUCTURE ONLY (see zk_oracle.h).
 *
 * Follows inflate/InflateStream.kt (whole file), inflate/StreamingBitReader.kt
 * (whole file) and CanonicalHuffman.decodeOne (inflate/CanonicalHuffman.kt:123-138).
 * The reference's three exception kinds are modelled with one longjmp target:
 *   SourceExhausted       -> Z_BUF_ERROR,  Pair.second = bytes out so far (IS:425-428)
 *   DataFormatException   -> Z_DATA_ERROR, Pair.second = bytes out so far (IS:429-432)
 *   IllegalStateException that is NOT wrapped (IS:243 -> CH:127,133) escapes the
 *   function in the reference -> ZKO_ESCAPED_EXCEPTION here.
 */
#include <setjmp.h>
#include <stdlib.h>
#include <string.h>

#include "zk_oracle.h"

enum { kWindow = 32768, kThrowExhausted = 1, kThrowFormat = 2, kThrowIllegalState = 3, kThrowSinkFull = 4 };

/* common/Constants.kt:237-316 */
static const int kLenBase[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27,
                                 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const int kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2,
                                  2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const int kDistBase[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,
                                  33,  49,  65,  97,  129, 193,  257,  385,  513,  769,
                                  1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const int kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6,
                                   6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const int kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

typedef struct {
    int max_len;
    int *bits; /* 2^max_len entries */
    int *vals;
} full_table;

typedef struct {
    /* StreamingBitReader state, BR:12-13 */
    const uint8_t *src;
    size_t src_len, src_pos;
    uint32_t hold;
    int held;
    /* sink + sliding window + running checksum, IS:386-389 */
    uint8_t *out;
    size_t out_cap, out_len;
    uint8_t window[kWindow];
    int wpos;
    uint32_t s1, s2;
    uint64_t out_count;
    jmp_buf thrown;
} inflater;

/* BR:48-57 */
static void br_need(inflater *z, int nbits) {
    while (z->held < nbits && z->src_pos < z->src_len) {
        z->hold |= (uint32_t)z->src[z->src_pos++] << z->held;
        z->held += 8;
    }
    if (nbits > 0 && z->held < nbits) longjmp(z->thrown, kThrowExhausted);
}
/* BR:20-25 */
static int br_peek(inflater *z, int nbits) {
    br_need(z, nbits);
    return nbits == 0 ? 0 : (int)(z->hold & ((1u << nbits) - 1u));
}
/* BR:27-32 */
static int br_take(inflater *z, int nbits) {
    int v = br_peek(z, nbits);
    z->hold >>= nbits;
    z->held -= nbits;
    return v;
}
/* BR:15-18 */
static void br_align(inflater *z) {
    int drop = z->held % 8;
    if (drop) br_take(z, drop);
}
/* BR:35-39 */
static int br_byte(inflater *z) {
    br_align(z);
    br_need(z, 8);
    return br_take(z, 8);
}

/* IS:81-98 -- one byte to sink + window, checksum reduced per byte */
static void put_byte(inflater *z, int b) {
    if (z->out_len >= z->out_cap) longjmp(z->thrown, kThrowSinkFull);
    z->out[z->out_len++] = (uint8_t)b;
    z->window[z->wpos] = (uint8_t)b;
    z->wpos = (z->wpos + 1) & (kWindow - 1);
    z->s1 = (z->s1 + (uint32_t)(b & 0xFF)) % 65521u;
    z->s2 = (z->s2 + z->s1) % 65521u;
    z->out_count++;
}

/* CH:123-138.  wrap != 0: caller is decodeSymOrThrow (IS:178-187), which turns the
 * IllegalStateException into a DataFormatException. */
static int decode_symbol(inflater *z, const full_table *t, int wrap) {
    if (t->max_len == 0) longjmp(z->thrown, wrap ? kThrowFormat : kThrowIllegalState);
    int look = br_peek(z, t->max_len);
    int len = t->bits[look];
    if (len == 0) longjmp(z->thrown, wrap ? kThrowFormat : kThrowIllegalState);
    int sym = t->vals[look];
    br_take(z, len);
    return sym;
}

static void table_alloc(full_table *t) {
    t->max_len = 0;
    t->bits = (int *)malloc(sizeof(int) * 32768);
    t->vals = (int *)malloc(sizeof(int) * 32768);
}
static int table_build(full_table *t, const int *lengths, int n) {
    t->max_len = zko_build_full(lengths, n, t->bits, t->vals);
    return t->max_len;
}
static void table_free(full_table *t) {
    free(t->bits);
    free(t->vals);
    t->bits = t->vals = NULL;
}

/* IS:47-74 */
static int stored_block(inflater *z) {
    br_align(z);
    int len = br_byte(z);
    len |= br_byte(z) << 8;
    int nlen = br_byte(z);
    nlen |= br_byte(z) << 8;
    if (((~len) & 0xFFFF) != nlen) return ZKO_DATA_ERROR; /* IS:58 */
    for (int i = 0; i < len; ++i) put_byte(z, br_byte(z));
    return ZKO_OK;
}

/* IS:132-152: "left" walk.  0 = fine, 1 = over-subscribed / incomplete. */
static int lengths_rejected(const int *lengths, int n) {
    int max_len = 0, used = 0;
    int per_len[16] = {0};
    for (int i = 0; i < n; ++i) {
        if (lengths[i] > max_len) max_len = lengths[i];
        if (lengths[i] > 0) {
            ++used;
            per_len[lengths[i]]++;
        }
    }
    if (max_len == 0) return 1;
    int left = 1;
    for (int b = 1; b <= max_len; ++b) {
        left = (left << 1) - per_len[b];
        if (left < 0) return 1;
    }
    return (left != 0 && used > 1) ? 1 : 0; /* IS:151: a single code may be incomplete */
}

/* IS:235-263.  Returns 0 on success, 1 where the reference returns null. */
static int read_code_lengths(inflater *z, const full_table *cl, int count, int *dst) {
    int i = 0;
    memset(dst, 0, sizeof(int) * (size_t)count);
    while (i < count) {
        int sym = decode_symbol(z, cl, /*wrap=*/0); /* IS:243 is NOT wrapped */
        if (sym <= 15) {
            dst[i++] = sym;
        } else if (sym == 16) {
            if (i == 0) return 1;
            int rep = 3 + br_take(z, 2), prev = dst[i - 1];
            while (rep--) {
                if (i >= count) return 1;
                dst[i++] = prev;
            }
        } else if (sym == 17 || sym == 18) {
            int rep = sym == 17 ? 3 + br_take(z, 3) : 11 + br_take(z, 7);
            while (rep--) {
                if (i >= count) return 1;
                dst[i++] = 0;
            }
        } else {
            return 1;
        }
    }
    return 0;
}

/* Shared symbol loop of IS:276-295 and IS:350-369. */
static int huffman_block(inflater *z, const full_table *lit, const full_table *dist) {
    for (;;) {
        int sym = decode_symbol(z, lit, 1);
        if (sym < 256) {
            put_byte(z, sym);
            continue;
        }
        if (sym == 256) return ZKO_OK;
        int lc = sym - 257;
        if (lc < 0 || lc > 28) return ZKO_DATA_ERROR; /* IS:195 */
        int length = kLenBase[lc] + (kLenExtra[lc] > 0 ? br_take(z, kLenExtra[lc]) : 0);
        int ds = decode_symbol(z, dist, 1);
        if (ds < 0 || ds > 29) return ZKO_DATA_ERROR; /* IS:207 */
        int distance = kDistBase[ds] + (kDistExtra[ds] > 0 ? br_take(z, kDistExtra[ds]) : 0);
        if (distance <= 0) return ZKO_DATA_ERROR;
        uint64_t reach = z->out_count < (uint64_t)kWindow ? z->out_count : (uint64_t)kWindow;
        if (distance < 1 || (uint64_t)distance > reach) return ZKO_DATA_ERROR; /* IS:287-291,361-365 */
        for (int k = 0; k < length; ++k) /* IS:216-232 */
            put_byte(z, z->window[(z->wpos - distance + kWindow) & (kWindow - 1)]);
    }
}

/* IS:154-173,265-297 */
static int fixed_block(inflater *z) {
    int ll[288], dl[32];
    for (int i = 0; i < 288; ++i) ll[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
    for (int i = 0; i < 32; ++i) dl[i] = 5;
    full_table lit, dist;
    table_alloc(&lit);
    table_alloc(&dist);
    table_build(&lit, ll, 288);
    table_build(&dist, dl, 32);
    /* decode under a nested jump target so the tables are released on every exit */
    jmp_buf outer;
    memcpy(&outer, &z->thrown, sizeof(jmp_buf));
    int thrown = setjmp(z->thrown);
    int rc = ZKO_OK;
    if (thrown == 0) rc = huffman_block(z, &lit, &dist);
    memcpy(&z->thrown, &outer, sizeof(jmp_buf));
    table_free(&lit);
    table_free(&dist);
    if (thrown) longjmp(z->thrown, thrown);
    return rc;
}

/* IS:299-371 */
static int dynamic_block(inflater *z) {
    int hlit = br_take(z, 5) + 257;
    int hdist = br_take(z, 5) + 1;
    int hclen = br_take(z, 4) + 4;
    int clen[19] = {0};
    for (int i = 0; i < hclen; ++i) clen[kClOrder[i]] = br_take(z, 3);

    /* storage is allocated before the jump target is armed so that no pointer
     * changes between setjmp and a longjmp back into this frame */
    full_table cl, lit, dist;
    table_alloc(&cl);
    table_alloc(&lit);
    table_alloc(&dist);
    int ll[320], dl[64];
    jmp_buf outer;
    memcpy(&outer, &z->thrown, sizeof(jmp_buf));
    volatile int rc = ZKO_OK;
    int thrown = setjmp(z->thrown);
    if (thrown == 0) {
        table_build(&cl, clen, 19); /* IS:315 -- built without validation */
        if (read_code_lengths(z, &cl, hlit, ll) || read_code_lengths(z, &cl, hdist, dl)) {
            rc = ZKO_DATA_ERROR; /* IS:316-317: two separate sequences */
        } else {
            int any = 0;
            for (int i = 0; i < hlit; ++i) any |= ll[i] != 0;
            int dany = 0;
            for (int i = 0; i < hdist; ++i) dany |= dl[i] != 0;
            if (!any || !(hlit > 256 && ll[256] > 0)) rc = ZKO_DATA_ERROR; /* IS:102-112,320-323 */
            else if (!dany) rc = ZKO_DATA_ERROR;                            /* IS:115-119,324-327 */
            else if (lengths_rejected(ll, hlit) || lengths_rejected(dl, hdist)) rc = ZKO_DATA_ERROR;
            else if (table_build(&lit, ll, hlit) <= 0 || table_build(&dist, dl, hdist) <= 0)
                rc = ZKO_DATA_ERROR; /* IS:122-128,337-348 */
            else rc = huffman_block(z, &lit, &dist);
        }
    }
    memcpy(&z->thrown, &outer, sizeof(jmp_buf));
    table_free(&cl);
    table_free(&lit);
    table_free(&dist);
    if (thrown) longjmp(z->thrown, thrown);
    return rc;
}

int zko_inflate_zlib(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap, size_t *out_len,
                     uint64_t *reported, size_t *consumed) {
    inflater *z = (inflater *)calloc(1, sizeof(inflater));
    z->src = in;
    z->src_len = in_len;
    z->out = out;
    z->out_cap = out_cap;
    z->s1 = 1; /* IS:388 */
    volatile uint64_t total_out = 0;
    volatile int rc = ZKO_OK;
    volatile uint64_t second = 0;

    int thrown = setjmp(z->thrown);
    if (thrown == 0) {
        /* IS:28-45 */
        br_align(z);
        int cmf = br_byte(z), flg = br_byte(z);
        if ((cmf & 0x0F) != 8 || ((cmf >> 4) & 0x0F) > 7 || ((cmf << 8) | flg) % 31 != 0) {
            rc = ZKO_DATA_ERROR;
        } else if (flg & 0x20) {
            for (int i = 0; i < 4; ++i) br_byte(z); /* IS:41: DICTID is consumed */
            rc = ZKO_NEED_DICT;
        }
        if (rc != ZKO_OK) { /* IS:384: (hdr, 0) */
            second = 0;
        } else {
            for (;;) { /* IS:391-411 */
                int last = br_take(z, 1);
                int type = br_take(z, 2);
                int r = type == 0 ? stored_block(z) : type == 1 ? fixed_block(z)
                        : type == 2 ? dynamic_block(z) : ZKO_DATA_ERROR;
                if (r != ZKO_OK) {
                    rc = r;
                    second = total_out; /* count as of the last completed block */
                    break;
                }
                total_out = z->out_count;
                if (last) break;
            }
            if (rc == ZKO_OK) { /* IS:413-424 */
                br_align(z);
                uint32_t trailer = 0;
                for (int i = 0; i < 4; ++i) trailer = (trailer << 8) | (uint32_t)br_byte(z);
                uint32_t running = (z->s2 << 16) | z->s1;
                rc = running == trailer ? ZKO_STREAM_END : ZKO_DATA_ERROR;
                second = total_out;
            }
        }
    } else {
        second = z->out_count;
        rc = thrown == kThrowExhausted ? ZKO_BUF_ERROR
             : thrown == kThrowFormat  ? ZKO_DATA_ERROR
             : thrown == kThrowSinkFull ? ZKO_MEM_ERROR
                                        : ZKO_ESCAPED_EXCEPTION;
    }
    if (out_len) *out_len = z->out_len;
    if (reported) *reported = second;
    if (consumed) *consumed = z->src_pos;
    int ret = rc;
    free(z);
    return ret;
}
