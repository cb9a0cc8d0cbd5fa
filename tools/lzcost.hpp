// lzcost.hpp -- deflate cost functions shared by the design-time CPU models (tools/lzmodel*.cpp): symbol mapping,
// zlib-style length-limited Huffman code lengths, dynamic-header size.  Not part of the product.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

static const int kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const int kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const int kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const int kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

static int len_sym(int len) {
    if (len == 258) return 28;
    int i = 0;
    while (i < 28 && kLenBase[i + 1] <= len) ++i;
    return i;
}
static int dist_sym(int d) {
    int i = 0;
    while (i < 29 && kDistBase[i + 1] <= d) ++i;
    return i;
}

// zlib-style length-limited Huffman: build the tree, then repair overflow on the length counts.
static void huff_lengths(const uint32_t *freq, int n, int maxbits, uint8_t *len) {
    std::vector<int> syms;
    for (int i = 0; i < n; ++i) {
        len[i] = 0;
        if (freq[i]) syms.push_back(i);
    }
    // force at least two codes so the code is complete (what zlib does)
    for (int i = 0; (int)syms.size() < 2 && i < n; ++i)
        if (!freq[i]) syms.push_back(i);
    std::vector<uint64_t> w;
    std::sort(syms.begin(), syms.end(), [&](int a, int b) {
        uint32_t fa = std::max(freq[a], 1u), fb = std::max(freq[b], 1u);
        return fa != fb ? fa < fb : a < b;
    });
    int m = (int)syms.size();
    std::vector<uint64_t> wt(2 * m);
    std::vector<int> parent(2 * m, -1);
    for (int i = 0; i < m; ++i) wt[i] = std::max(freq[syms[i]], 1u);
    int leaf = 0, inner = m, next = m;
    auto take = [&]() {
        if (leaf < m && (inner >= next || wt[leaf] <= wt[inner])) return leaf++;
        return inner++;
    };
    while (next < 2 * m - 1) {
        int a = take(), b = take();
        wt[next] = wt[a] + wt[b];
        parent[a] = parent[b] = next;
        ++next;
    }
    std::vector<int> depth(2 * m, 0);
    for (int i = 2 * m - 3; i >= 0; --i) depth[i] = depth[parent[i]] + 1;
    std::vector<int> count(64, 0);
    int overflow = 0;
    for (int i = 0; i < m; ++i) {
        int d = depth[i];
        if (d > maxbits) {
            d = maxbits;
            ++overflow;
        }
        count[d]++;
    }
    while (overflow > 0) {
        int bits = maxbits - 1;
        while (count[bits] == 0) --bits;
        count[bits]--;
        count[bits + 1] += 2;
        count[maxbits]--;
        overflow -= 2;
    }
    // longest codes go to the least frequent symbols (syms is ascending by weight)
    int idx = 0;
    for (int bits = maxbits; bits >= 1; --bits)
        for (int k = 0; k < count[bits]; ++k) len[syms[idx++]] = (uint8_t)bits;
}

static uint64_t header_bits(const uint8_t *ll, const uint8_t *dl) {
    int hlit = 286, hdist = 30;
    while (hlit > 257 && ll[hlit - 1] == 0) --hlit;
    while (hdist > 1 && dl[hdist - 1] == 0) --hdist;
    uint32_t clf[19] = {0};
    uint64_t extra = 0;
    auto rle = [&](const uint8_t *l, int n) {
        int i = 0;
        while (i < n) {
            int v = l[i], j = i;
            while (j < n && l[j] == v) ++j;
            int run = j - i;
            i = j;
            if (v == 0) {
                while (run >= 11) { int r = std::min(run, 138); clf[18]++; extra += 7; run -= r; }
                if (run >= 3) { clf[17]++; extra += 3; run = 0; }
                clf[0] += run;
            } else {
                clf[v]++; --run;
                while (run >= 3) { int r = std::min(run, 6); clf[16]++; extra += 2; run -= r; }
                clf[v] += run;
            }
        }
    };
    rle(ll, hlit);
    rle(dl, hdist);
    uint8_t bl[19];
    huff_lengths(clf, 19, 7, bl);
    static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int hclen = 19;
    while (hclen > 4 && bl[order[hclen - 1]] == 0) --hclen;
    uint64_t bits = 3 + 5 + 5 + 4 + 3 * hclen + extra;
    for (int i = 0; i < 19; ++i) bits += (uint64_t)clf[i] * bl[i];
    return bits;
}

