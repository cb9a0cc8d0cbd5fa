/*
 * ref_deflate.c -- oracle restatement of DeflateStream.compressZlib, quirks included.
 * TEST INFRASTRUCTURE ONLY (see zk_oracle.h).
 *
 * Follows deflate/DeflateStream.kt (DS) and inflate/StreamingBitWriter.kt (BW).
 * This is NOT a good compressor and is not meant to be: it exists to state the
 * reference's compressed SIZE at a level and to produce reference streams for
 * the GPU inflater.  Behaviour kept on purpose:
 *   - every look-ahead position is inserted into the hash chains before the
 *     tokeniser runs, so chains start with FUTURE positions that are skipped but
 *     still spend the chain budget (DS:654-658, 672-674, 689-690);
 *   - the insert guard compares an index with a length (DS:138);
 *   - inserts issued while a match is being stepped over see the tokeniser's
 *     stale pos/laOff/laLen (DS:597-599 captured by DS:165-184);
 *   - candidate bytes are read from `window`, which only holds bytes < pos, so an
 *     overlapping candidate is compared against stale bytes (DS:677) -- streams
 *     produced from ordinary text can therefore decode to different data;
 *   - length 258 is coded as symbol 284 + 5 extra bits (DS:193-203);
 *   - code lengths are Shannon-Fano and may be over-subscribed at the depth cap
 *     (HuffmanBuilder.kt:58-62);
 *   - the level<=0 "stored" path reads every Okio segment to offset 0 of its
 *     buffer (DS:350), so blocks longer than one 8 KiB segment carry stale bytes;
 *   - lazy matching is compiled out (DS:610).
 * Okio model: BufferedSource.read(ByteArray, off, n) hands out at most the rest
 * of the current 8192-byte segment (okio 3.10.2 Buffer.read / RealBufferedSource.read);
 * no reference test pins that granularity.
 */
#include <stdlib.h>
#include <string.h>

#include "zk_oracle.h"

enum { kWin = 32768, kMask = kWin - 1, kMaxStored = 65535, kSegment = 8192 };

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

/* ---- Okio stand-ins ---------------------------------------------------- */
typedef struct {
    const uint8_t *data;
    size_t n, at;
} source_t;

static int src_exhausted(const source_t *s) { return s->at >= s->n; }
/* at most the remainder of the current 8 KiB segment; -1 at end of input */
static int src_read(source_t *s, uint8_t *dst, int want) {
    if (s->at >= s->n) return -1;
    size_t seg_left = kSegment - (s->at % kSegment);
    size_t k = s->n - s->at;
    if (k > seg_left) k = seg_left;
    if (k > (size_t)want) k = (size_t)want;
    memcpy(dst, s->data + s->at, k);
    s->at += k;
    return (int)k;
}

typedef struct {
    uint8_t *p;
    size_t n, cap;
} sink_t;

static void sink_byte(sink_t *k, int b) {
    if (k->n == k->cap) {
        k->cap = k->cap ? k->cap * 2 : 4096;
        k->p = (uint8_t *)realloc(k->p, k->cap);
    }
    k->p[k->n++] = (uint8_t)b;
}
static void sink_bytes(sink_t *k, const uint8_t *b, size_t n) {
    for (size_t i = 0; i < n; ++i) sink_byte(k, b[i]);
}

/* ---- StreamingBitWriter, BW:10-49 -------------------------------------- */
typedef struct {
    sink_t *sink;
    uint32_t acc;
    int fill;
} bitw_t;

static void bw_put(bitw_t *w, int value, int count) { /* BW:18-35 */
    if (count == 0) return;
    uint32_t v = (uint32_t)value & ((1u << count) - 1u);
    w->acc |= v << w->fill;
    w->fill += count;
    while (w->fill >= 8) {
        sink_byte(w->sink, (int)(w->acc & 0xFF));
        w->acc >>= 8;
        w->fill -= 8;
    }
}
static void bw_align(bitw_t *w) { /* BW:37-40 */
    int rem = w->fill % 8;
    if (rem) bw_put(w, 0, 8 - rem);
}

/* ---- zlib framing, DS:27-95 -------------------------------------------- */
static void put_zlib_header(sink_t *k, int level) {
    int flevel = level >= 9 ? 3 : level >= 6 ? 2 : level >= 2 ? 1 : 0; /* DS:27-33 */
    int cmf = 0x78, flg = flevel << 6;
    int fcheck = (31 - (((cmf << 8) | flg) % 31)) % 31; /* DS:45-47 */
    sink_byte(k, cmf);
    sink_byte(k, (flg & 0xE0) | fcheck);
}
static void put_be32(sink_t *k, uint32_t v) {
    for (int sh = 24; sh >= 0; sh -= 8) sink_byte(k, (int)((v >> sh) & 0xFF));
}
static void put_stored_header(bitw_t *w, int length, int bfinal) { /* DS:80-95 */
    bw_put(w, bfinal, 1);
    bw_put(w, 0, 2);
    bw_align(w);
    int len = length & 0xFFFF, nlen = (~len) & 0xFFFF;
    sink_byte(w->sink, len & 0xFF);
    sink_byte(w->sink, len >> 8);
    sink_byte(w->sink, nlen & 0xFF);
    sink_byte(w->sink, nlen >> 8);
}

/* ---- symbol mapping by linear scan, DS:187-228 -------------------------- */
typedef struct {
    int sym, nextra, extra;
} mapped_t;

static mapped_t map_length(int length) {
    mapped_t m = {-1, 0, 0};
    for (int i = 0; i < 29; ++i) {
        int top = kLenBase[i] + (kLenExtra[i] > 0 ? (1 << kLenExtra[i]) - 1 : 0);
        if (length >= kLenBase[i] && length <= top) { /* first hit wins: 258 lands on i=27 */
            m.sym = 257 + i;
            m.nextra = kLenExtra[i];
            m.extra = length - kLenBase[i];
            break;
        }
    }
    return m;
}
static mapped_t map_dist(int dist) {
    mapped_t m = {-1, 0, 0};
    for (int i = 0; i < 30; ++i) {
        int top = kDistBase[i] + (kDistExtra[i] > 0 ? (1 << kDistExtra[i]) - 1 : 0);
        if (dist >= kDistBase[i] && dist <= top) {
            m.sym = i;
            m.nextra = kDistExtra[i];
            m.extra = dist - kDistBase[i];
            break;
        }
    }
    return m;
}

/* ---- LZ77 state shared by the fixed and dynamic paths ------------------- */
typedef struct {
    uint8_t la[kWin];     /* look-ahead buffer (DS:406,575) */
    uint8_t window[kWin]; /* bytes already stepped over (DS:407,576) */
    int head[kWin], prev[kWin];
    int pos, la_off, la_len;
} lz_t;

static int hash3(const uint8_t *p) { /* DS:419-427,588-595 */
    return (p[0] * 251 + p[1] * 271 + p[2] * 277) & (kWin - 1);
}

/* DS:127-144.  view_* are the tokeniser variables as the closure at DS:429-431 /
 * DS:597-599 sees them at call time. */
static void insert_triplet(lz_t *z, int abs_pos, int view_pos, int view_off, int view_len) {
    int rel = abs_pos - view_pos + view_off;
    if (view_len - rel < 3) return; /* DS:138 (index used as a length) */
    int h = hash3(z->la + rel);
    z->prev[abs_pos & kMask] = z->head[h];
    z->head[h] = abs_pos;
}

/* DS:165-184 driven as at DS:503-509 / DS:746-752: the inserts see the values
 * pos/laOff/laLen had BEFORE the match is stepped over. */
static void step_over_match(lz_t *z, int match_len, int available) {
    int pos0 = z->pos, off0 = z->la_off, len0 = z->la_len;
    int pos = pos0, off = off0, len = len0;
    for (int k = 0; k < match_len; ++k) {
        z->window[pos & kMask] = z->la[off];
        if (available - k >= 3) insert_triplet(z, pos, pos0, off0, len0);
        ++pos;
        ++off;
        --len;
    }
    z->pos = pos;
    z->la_off = off;
    z->la_len = len;
}

static void step_over_literal(lz_t *z, int available) { /* DS:512-517, 756-761 */
    z->window[z->pos & kMask] = z->la[z->la_off];
    if (available >= 3) insert_triplet(z, z->pos, z->pos, z->la_off, z->la_len);
    z->pos++;
    z->la_off++;
    z->la_len--;
}

/* DS:476-497 / DS:668-692: walk head/prev for at most max_chain entries. */
static void longest_match(const lz_t *z, int available, int max_chain, int *best_len, int *best_dist) {
    *best_len = 0;
    *best_dist = 0;
    int rel = z->la_off;
    if (available < 3 || rel + 2 >= kWin) return;
    int m = z->head[hash3(z->la + rel)];
    for (int chain = 0; m != -1 && chain < max_chain; ++chain) {
        int dist = z->pos - m;
        if (dist >= 1 && dist <= kWin) {
            int run = 0;
            while (run < 258 && run < available) {
                if (rel + run >= kWin) break;
                if (z->window[(m + run) & kMask] != z->la[rel + run]) break; /* stale bytes when m+run >= pos */
                ++run;
            }
            if (run >= 3 && run > *best_len) {
                *best_len = run;
                *best_dist = dist;
                if (run >= 258) break;
            }
        }
        m = z->prev[m & kMask];
    }
}

static void compact_lookahead(lz_t *z) { /* DS:523-528, 766-771 */
    if (z->la_off > 0 && z->la_len > 0) {
        memmove(z->la, z->la + z->la_off, (size_t)z->la_len);
        z->la_off = 0;
    } else if (z->la_len == 0) {
        z->la_off = 0;
    }
}

static void seed_range(lz_t *z, int from) { /* DS:460-464, 535-540, 654-658, 782-786 */
    for (int p = from; p + 2 < z->pos + z->la_len; ++p) insert_triplet(z, p, z->pos, z->la_off, z->la_len);
}

static lz_t *lz_new(void) {
    lz_t *z = (lz_t *)calloc(1, sizeof(lz_t));
    for (int i = 0; i < kWin; ++i) z->head[i] = z->prev[i] = -1;
    return z;
}

/* ---- level <= 0: stored, DS:331-373 -------------------------------------- */
static uint64_t deflate_stored(source_t *src, sink_t *out, int level) {
    put_zlib_header(out, level);
    bitw_t w = {out, 0, 0};
    uint8_t *buf = (uint8_t *)calloc(1, 64 * 1024);
    uint64_t total_in = 0;
    uint32_t adler = 1;
    int eof = 0;
    while (!eof) {
        int filled = 0;
        while (filled < kMaxStored && !src_exhausted(src)) {
            int want = kMaxStored - filled;
            if (want > 64 * 1024) want = 64 * 1024;
            int n = src_read(src, buf /* offset 0 every time, DS:350 */, want);
            if (n <= 0) break;
            filled += n;
        }
        if (filled < kMaxStored && src_exhausted(src)) eof = 1;
        put_stored_header(&w, filled, eof);
        if (filled > 0) {
            sink_bytes(out, buf, (size_t)filled);
            total_in += (uint64_t)filled;
            adler = zko_adler32(adler, buf, (size_t)filled);
        }
    }
    put_be32(out, adler);
    free(buf);
    return total_in;
}

/* ---- level == 1: one fixed-Huffman block, DS:376-553 ---------------------- */
static void fixed_lengths(int *ll /*288*/, int *dl /*32*/) { /* DS:285-299 */
    for (int i = 0; i < 288; ++i) ll[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
    for (int i = 0; i < 32; ++i) dl[i] = 5;
}

static void put_match(bitw_t *w, const int *lc, const int *lb, const int *dc, const int *db, int len,
                      int dist) { /* DS:231-252 with the <=258 split of DS:441-447 / DS:967-974 */
    int remaining = len;
    while (remaining > 0) {
        int l = remaining < 258 ? remaining : 258;
        mapped_t m = map_length(l);
        bw_put(w, lc[m.sym], lb[m.sym]);
        if (m.nextra > 0) bw_put(w, m.extra, m.nextra);
        mapped_t d = map_dist(dist);
        bw_put(w, dc[d.sym], db[d.sym]);
        if (d.nextra > 0) bw_put(w, d.extra, d.nextra);
        remaining -= l;
    }
}

static uint64_t deflate_fixed(source_t *src, sink_t *out, int level) {
    put_zlib_header(out, level);
    bitw_t w = {out, 0, 0};
    int ll[288], dl[32], lc[288], lb[288], dc[32], db[32];
    fixed_lengths(ll, dl);
    zko_build_encoder(ll, 288, lc, lb);
    zko_build_encoder(dl, 32, dc, db);
    bw_put(&w, 1, 1); /* DS:400-401 */
    bw_put(&w, 1, 2);

    lz_t *z = lz_new();
    uint32_t adler = 1;
    uint64_t total_in = 0;
    while (z->la_len < kWin && !src_exhausted(src)) { /* DS:451-457 */
        int n = src_read(src, z->la + z->la_len, kWin - z->la_len);
        if (n <= 0) break;
        adler = zko_adler32(adler, z->la + z->la_len, (size_t)n);
        z->la_len += n;
        total_in += (uint64_t)n;
    }
    seed_range(z, z->pos);
    while (z->la_len > 0) { /* DS:467-543 */
        int available = z->la_len, blen, bdist;
        int b0 = z->la[z->la_off];
        longest_match(z, available, 32, &blen, &bdist);
        if (blen >= 3) {
            put_match(&w, lc, lb, dc, db, blen, bdist);
            step_over_match(z, blen, available);
        } else {
            bw_put(&w, lc[b0], lb[b0]);
            step_over_literal(z, available);
        }
        if (z->la_len < 1024 && !src_exhausted(src)) { /* DS:521-542 */
            compact_lookahead(z);
            int n = src_read(src, z->la + z->la_len, kWin - z->la_len);
            if (n > 0) {
                adler = zko_adler32(adler, z->la + z->la_len, (size_t)n);
                z->la_len += n;
                total_in += (uint64_t)n;
                seed_range(z, z->pos + z->la_len - n);
            }
        }
    }
    bw_put(&w, lc[256], lb[256]); /* DS:546 */
    bw_align(&w);
    put_be32(out, adler);
    free(z);
    return total_in;
}

/* ---- level >= 2: per-block stored / fixed / dynamic choice, DS:556-1046 --- */
typedef struct {
    int sym, nextra, extra;
} clsym_t;

/* DS:834-878 */
static int rle_lengths(const int *lengths, int count, clsym_t *dst) {
    int n = 0, i = 0, prev = -1;
    while (i < count) {
        int l = lengths[i];
        if (l == 0) {
            int run = 1;
            while (i + run < count && lengths[i + run] == 0 && run < 138) ++run;
            if (run >= 11) dst[n++] = (clsym_t){18, 7, run - 11};
            else if (run >= 3) dst[n++] = (clsym_t){17, 3, run - 3};
            else
                for (int r = 0; r < run; ++r) dst[n++] = (clsym_t){0, 0, 0};
            i += run;
            prev = 0;
        } else {
            int run = 1;
            while (i + run < count && lengths[i + run] == l && run < 6) ++run;
            if (prev == l && run >= 3) {
                dst[n++] = (clsym_t){16, 2, run - 3};
            } else {
                dst[n++] = (clsym_t){l, 0, 0};
                if (run >= 2) {
                    int rem = run - 1;
                    if (rem >= 3) dst[n++] = (clsym_t){16, 2, rem - 3};
                    else
                        for (int r = 0; r < rem; ++r) dst[n++] = (clsym_t){l, 0, 0};
                }
            }
            i += run;
            prev = l;
        }
    }
    return n;
}

typedef struct {
    uint16_t len;  /* 0 = literal */
    uint16_t aux;  /* literal byte, or distance-1 */
} token_t;

static int64_t token_bits(const token_t *tok, int ntok, const int *lbits, const int *dbits) { /* DS:902-923 */
    int64_t bits = 0;
    for (int i = 0; i < ntok; ++i) {
        if (tok[i].len == 0) {
            bits += lbits[tok[i].aux];
        } else {
            mapped_t m = map_length(tok[i].len), d = map_dist(tok[i].aux + 1);
            bits += lbits[m.sym] + m.nextra + dbits[d.sym] + d.nextra;
        }
    }
    return bits + lbits[256];
}

static void put_tokens(bitw_t *w, const token_t *tok, int ntok, const int *lc, const int *lb, const int *dc,
                       const int *db) { /* DS:958-1000 */
    for (int i = 0; i < ntok; ++i) {
        if (tok[i].len == 0) bw_put(w, lc[tok[i].aux], lb[tok[i].aux]);
        else put_match(w, lc, lb, dc, db, tok[i].len, tok[i].aux + 1);
    }
    bw_put(w, lc[256], lb[256]);
}

static uint64_t deflate_dynamic(source_t *src, sink_t *out, int level) {
    put_zlib_header(out, level);
    bitw_t w = {out, 0, 0};
    lz_t *z = lz_new();
    uint32_t adler = 1;
    uint64_t total_in = 0;
    int max_chain = level <= 2 ? 8 : level <= 4 ? 16 : level <= 6 ? 32 : 64; /* DS:603-609 */
    token_t *tok = (token_t *)malloc(sizeof(token_t) * (kMaxStored + 8));
    uint8_t *raw = (uint8_t *)malloc(kMaxStored);
    int fll[288], fdl[32], flc[288], flb[288], fdc[32], fdb[32];
    fixed_lengths(fll, fdl);
    zko_build_encoder(fll, 288, flc, flb);
    zko_build_encoder(fdl, 32, fdc, fdb);

    for (;;) {
        if (z->la_off > 0) { /* DS:614-623 */
            if (z->la_len > 0) memmove(z->la, z->la + z->la_off, (size_t)z->la_len);
            z->la_off = 0;
        }
        int ntok = 0, lit_freq[286] = {0}, dist_freq[30] = {0};
        int block_read = 0, raw_len = 0;

        while (z->la_len < kWin && block_read < kMaxStored && !src_exhausted(src)) { /* DS:641-651 */
            int want = kWin - z->la_len;
            if (want > kMaxStored - block_read) want = kMaxStored - block_read;
            int n = src_read(src, z->la + z->la_len, want);
            if (n <= 0) break;
            if (raw_len + n <= kMaxStored) { /* DS:633-638 */
                memcpy(raw + raw_len, z->la + z->la_len, (size_t)n);
                raw_len += n;
            }
            adler = zko_adler32(adler, z->la + z->la_len, (size_t)n);
            z->la_len += n;
            total_in += (uint64_t)n;
            block_read += n;
        }
        seed_range(z, z->pos); /* DS:654-658 */

        while (z->la_len > 0) { /* DS:661-790 */
            int rel = z->la_off;
            int available = z->la_len < kWin - z->la_off ? z->la_len : kWin - z->la_off;
            if (rel >= kWin) break;
            int b0 = z->la[rel], blen, bdist;
            longest_match(z, available, max_chain, &blen, &bdist);
            if (blen >= 3) { /* DS:735-752 (doLazy == false, DS:610) */
                tok[ntok++] = (token_t){(uint16_t)blen, (uint16_t)(bdist - 1)};
                lit_freq[map_length(blen).sym]++;
                dist_freq[map_dist(bdist).sym]++;
                step_over_match(z, blen, available);
            } else { /* DS:753-762 */
                tok[ntok++] = (token_t){0, (uint16_t)b0};
                lit_freq[b0]++;
                step_over_literal(z, available);
            }
            if (z->la_len < 1024 && block_read < kMaxStored && !src_exhausted(src)) { /* DS:765-789 */
                compact_lookahead(z);
                int want = kWin - z->la_len;
                if (want > kMaxStored - block_read) want = kMaxStored - block_read;
                if (want > 0) {
                    int n = src_read(src, z->la + z->la_len, want);
                    if (n > 0) {
                        if (raw_len + n <= kMaxStored) {
                            memcpy(raw + raw_len, z->la + z->la_len, (size_t)n);
                            raw_len += n;
                        }
                        adler = zko_adler32(adler, z->la + z->la_len, (size_t)n);
                        z->la_len += n;
                        total_in += (uint64_t)n;
                        block_read += n;
                        seed_range(z, z->pos + z->la_len - n);
                    }
                }
            }
        }

        int is_last = src_exhausted(src) && z->la_len == 0; /* DS:793 */
        lit_freq[256]++;                                    /* DS:796-801 */
        int any_dist = 0;
        for (int i = 0; i < 30; ++i) any_dist |= dist_freq[i] != 0;
        if (!any_dist) dist_freq[0] = 1; /* DS:803-812 */

        int dll[286], ddl[30];
        zko_build_lengths(lit_freq, 286, 15, 256, dll); /* DS:815-816 */
        zko_build_lengths(dist_freq, 30, 15, 0, ddl);
        int last_lit = 285, last_dist = 29;
        while (last_lit >= 0 && dll[last_lit] == 0) --last_lit;
        while (last_dist >= 0 && ddl[last_dist] == 0) --last_dist;
        if (last_lit < 256) last_lit = 256; /* DS:823-824 */
        if (last_dist < 0) last_dist = 0;

        clsym_t cl[320];
        int ncl = rle_lengths(dll, last_lit + 1, cl); /* DS:880-886: two runs, never merged */
        ncl += rle_lengths(ddl, last_dist + 1, cl + ncl);
        int cl_freq[19] = {0}, bll[19];
        for (int i = 0; i < ncl; ++i) cl_freq[cl[i].sym]++;
        zko_build_lengths(cl_freq, 19, 7, 0, bll); /* DS:889-891 */
        int hclen = 19;
        while (hclen > 4 && bll[kClOrder[hclen - 1]] == 0) --hclen; /* DS:893-894 */

        int dlc[286], dlb[286], ddc[30], ddb[30], blc[19], blb[19];
        zko_build_encoder(dll, 286, dlc, dlb); /* DS:898-900 */
        zko_build_encoder(ddl, 30, ddc, ddb);
        zko_build_encoder(bll, 19, blc, blb);

        int64_t head_bits = 1 + 2 + 5 + 5 + 4 + 3 * (int64_t)hclen; /* DS:925-935 */
        for (int i = 0; i < ncl; ++i) head_bits += blb[cl[i].sym] + cl[i].nextra;
        int64_t cost_dyn = head_bits + token_bits(tok, ntok, dlb, ddb);
        int64_t cost_fix = 1 + 2 + token_bits(tok, ntok, flb, fdb);        /* DS:944 */
        int pad = (8 - (w.fill % 8)) % 8;                                  /* DS:946 */
        int64_t cost_sto = pad + 1 + 2 + 16 + 16 + ((int64_t)raw_len << 3); /* DS:947 */

        if (cost_sto <= cost_dyn && cost_sto <= cost_fix) { /* DS:950-955 */
            put_stored_header(&w, raw_len, is_last);
            if (raw_len > 0) sink_bytes(out, raw, (size_t)raw_len);
        } else if (cost_dyn <= cost_fix) { /* DS:1008-1029 */
            bw_put(&w, is_last, 1);
            bw_put(&w, 2, 2);
            bw_put(&w, last_lit + 1 - 257, 5);
            bw_put(&w, last_dist, 5);
            bw_put(&w, hclen - 4, 4);
            for (int i = 0; i < hclen; ++i) bw_put(&w, bll[kClOrder[i]], 3);
            for (int i = 0; i < ncl; ++i) {
                bw_put(&w, blc[cl[i].sym], blb[cl[i].sym]);
                if (cl[i].nextra) bw_put(&w, cl[i].extra, cl[i].nextra);
            }
            put_tokens(&w, tok, ntok, dlc, dlb, ddc, ddb);
        } else { /* DS:1030-1035 */
            bw_put(&w, is_last, 1);
            bw_put(&w, 1, 2);
            put_tokens(&w, tok, ntok, flc, flb, fdc, fdb);
        }
        if (is_last) break;
    }
    bw_align(&w); /* DS:1040 */
    put_be32(out, adler);
    free(raw);
    free(tok);
    free(z);
    return total_in;
}

int zko_deflate_zlib(const uint8_t *in, size_t n, int level, uint8_t **out, size_t *out_len, uint64_t *bytes_in) {
    *out = NULL;
    *out_len = 0;
    if (bytes_in) *bytes_in = 0;
    if (level > 9) return ZKO_STREAM_ERROR; /* DS:323-325: nothing is written */
    source_t src = {in, n, 0};
    sink_t sink = {NULL, 0, 0};
    uint64_t consumed = level <= 0   ? deflate_stored(&src, &sink, level) /* DS:307-311 */
                        : level == 1 ? deflate_fixed(&src, &sink, level)
                                     : deflate_dynamic(&src, &sink, level);
    *out = sink.p;
    *out_len = sink.n;
    if (bytes_in) *bytes_in = consumed;
    return ZKO_OK;
}

void zko_free(void *p) { free(p); }
