/*
 * ref_huffman.c -- oracle restatement of the reference's code-length builder
 * and canonical-code helpers.  TEST INFRASTRUCTURE ONLY (see zk_oracle.h).
 *
 *   zko_build_lengths : deflate/HuffmanBuilder.kt:10-83   (Shannon-Fano split, NOT Huffman)
 *   zko_build_encoder : inflate/CanonicalHuffman.kt:85-120,141-154
 *   zko_build_full    : inflate/CanonicalHuffman.kt:54-82
 *
 * The reference routes shifts through ArithmeticBitwiseOps (ABO:62-99,189-226);
 * for the operand ranges used here those are ordinary 32-bit shifts / ors.
 */
#include <stdlib.h>
#include <string.h>

#include "zk_oracle.h"

typedef struct {
    int sym;
    int64_t weight;
} sf_item;

typedef struct {
    const sf_item *items;
    const int64_t *cum; /* cum[i] = sum of weights of items[0..i) */
    int *lengths;
    int max_bits;
} sf_ctx;

/* HB:43 -- descending frequency, ties by ascending symbol id */
static int sf_order(const void *pa, const void *pb) {
    const sf_item *a = (const sf_item *)pa, *b = (const sf_item *)pb;
    if (a->weight != b->weight) return a->weight > b->weight ? -1 : 1;
    return a->sym - b->sym;
}

/* HB:48-79 -- recursive halving of [lo,hi) at the prefix closest to half the weight */
static void sf_split(const sf_ctx *c, int lo, int hi, int depth) {
    if (lo >= hi) return;
    if (hi - lo == 1) { /* HB:54-57 */
        c->lengths[c->items[lo].sym] = depth > 1 ? depth : 1;
        return;
    }
    if (depth >= c->max_bits) { /* HB:58-62 -- may over-subscribe the code; kept as is */
        for (int i = lo; i < hi; ++i) c->lengths[c->items[i].sym] = c->max_bits;
        return;
    }
    int64_t total = c->cum[hi] - c->cum[lo];
    int64_t half = total / 2;
    int cut = lo + 1;
    int64_t best = INT64_MAX;
    for (int i = lo + 1; i < hi; ++i) { /* HB:63-75: first minimum wins (strict <) */
        int64_t left = c->cum[i] - c->cum[lo];
        int64_t gap = left > half ? left - half : half - left;
        if (gap < best) {
            best = gap;
            cut = i;
        }
    }
    sf_split(c, lo, cut, depth + 1);
    sf_split(c, cut, hi, depth + 1);
}

void zko_build_lengths(const int *freq, int n, int max_bits, int ensure_symbol, int *out_lengths) {
    memset(out_lengths, 0, sizeof(int) * (size_t)n);
    /* HB:17 -- totalFreq is an Int sum in the reference; block sizes keep it far from overflow */
    int total = 0;
    for (int i = 0; i < n; ++i) total += freq[i];
    if (total == 0) { /* HB:20-25 */
        if (ensure_symbol >= 0 && ensure_symbol < n) out_lengths[ensure_symbol] = 1;
        return;
    }
    sf_item *items = (sf_item *)malloc(sizeof(sf_item) * (size_t)n);
    int used = 0;
    for (int i = 0; i < n; ++i)
        if (freq[i] > 0) { /* HB:33-34 */
            items[used].sym = i;
            items[used].weight = freq[i];
            ++used;
        }
    if (used == 1) { /* HB:37-40 */
        out_lengths[items[0].sym] = 1;
        free(items);
        return;
    }
    qsort(items, (size_t)used, sizeof(sf_item), sf_order);
    int64_t *cum = (int64_t *)malloc(sizeof(int64_t) * (size_t)(used + 1));
    cum[0] = 0;
    for (int i = 0; i < used; ++i) cum[i + 1] = cum[i] + items[i].weight; /* HB:45-46 */
    sf_ctx c = {items, cum, out_lengths, max_bits};
    sf_split(&c, 0, used, 0); /* HB:81 */
    free(cum);
    free(items);
}

/* CH:141-154 */
static int mirror_bits(int value, int width) {
    int r = 0;
    for (int i = 0; i < width; ++i) {
        r = (r << 1) | (value & 1);
        value >>= 1;
    }
    return r;
}

/* CH:85-100 -- first canonical code of every length */
static void first_codes(const int *lengths, int n, int max_len, int *next_code /* [max_len+1] */) {
    int per_len[17];
    memset(per_len, 0, sizeof per_len);
    for (int i = 0; i < n; ++i)
        if (lengths[i] > 0) per_len[lengths[i]]++;
    int code = 0;
    next_code[0] = 0;
    for (int b = 1; b <= max_len; ++b) {
        code = (code + per_len[b - 1]) << 1;
        next_code[b] = code;
    }
}

static int max_length(const int *lengths, int n) {
    int m = 0;
    for (int i = 0; i < n; ++i)
        if (lengths[i] > m) m = lengths[i];
    return m;
}

void zko_build_encoder(const int *lengths, int n, int *codes, int *bits) { /* CH:103-120 */
    int max_len = max_length(lengths, n);
    int next_code[17];
    memset(next_code, 0, sizeof next_code);
    first_codes(lengths, n, max_len > 16 ? 16 : max_len, next_code);
    for (int s = 0; s < n; ++s) {
        codes[s] = 0;
        bits[s] = 0;
        int len = lengths[s];
        if (len == 0) continue;
        int assigned = next_code[len]++;
        codes[s] = mirror_bits(assigned, len);
        bits[s] = len;
    }
}

int zko_build_full(const int *lengths, int n, int *bits_tab, int *vals_tab) { /* CH:54-82 */
    int max_len = max_length(lengths, n);
    if (max_len == 0) { /* CH:57 -- FullTable(0, IntArray(1), IntArray(1)) */
        bits_tab[0] = 0;
        vals_tab[0] = 0;
        return 0;
    }
    if (max_len > 15) return -1; /* CH:22 require(maxLen <= 15) throws */
    int next_code[17];
    first_codes(lengths, n, max_len, next_code);
    int size = 1 << max_len;
    memset(bits_tab, 0, sizeof(int) * (size_t)size);
    memset(vals_tab, 0, sizeof(int) * (size_t)size);
    for (int s = 0; s < n; ++s) {
        int len = lengths[s];
        if (len == 0) continue;
        int assigned = next_code[len]++;
        int wire = mirror_bits(assigned, len);
        /* CH:74-79: every index whose low `len` bits equal the wire code.  Later
         * symbols overwrite earlier ones when the code is over-subscribed, and an
         * over-subscribed `assigned` can exceed `len` bits: mirror_bits then drops
         * the high bits exactly as the reference's reverseBits does. */
        for (int idx = wire; idx < size; idx += 1 << len) {
            bits_tab[idx] = len;
            vals_tab[idx] = s;
        }
    }
    return max_len;
}
