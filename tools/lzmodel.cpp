// lzmodel.cpp -- design-time CPU model of the GPU match-finder / parser, used to tune
// ratio parameters (hash width, chain depth, lazy rule, block span) against libz and the
// oracle.  Not part of the product and not a fallback: it only prints sizes.
//
//   g++ -O2 -o /tmp/lzmodel tools/lzmodel.cpp && /tmp/lzmodel file [key=value ...]
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct Params {
    int hb = 15;        // hash bits
    int hbytes = 4;     // bytes hashed (3 or 4)
    int chain = 16;     // max chain steps
    int nice = 64;      // stop searching at this length
    int lazy = 1;       // one-step lazy evaluation
    int run = 8;        // inheritance run length (positions handled by one lane in sequence)
    int block = 32768;  // deflate block span in input bytes
    int unit = 524288;  // LZ unit (history re-primed at unit start)
    int group = 2048;   // pipeline group; max distance = 32768 - 2*group
    int minlen = 4;
    int toofar = 4096;  // a length-3 match farther than this is dropped
    int good = 8;       // zlib-style: quarter the chain when the inherited/previous match is >= good
    int skip2 = 0;      // >0: from the 3rd position of an inherited match on, chain budget = chain >> skip2 (99 = none)
    int inhlen = 0;     // count of consecutive inherited positions
    int cap = 0;        // >0: candidates are compared up to `cap` bytes; the first (nearest) one reaching the cap wins and is extended
    int probe3 = 0;     // 1: when nothing >= 4 was found, try length-3+ matches at a few fixed small distances
    int lit4 = 0;       // >0: a length-4 match farther than this is dropped
};

#include "lzcost.hpp"

int main(int argc, char **argv) {
    if (argc < 2) return 1;
    Params P;
    for (int i = 2; i < argc; ++i) {
        char *eq = strchr(argv[i], '=');
        if (!eq) continue;
        std::string k(argv[i], eq - argv[i]);
        int v = atoi(eq + 1);
        if (k == "hb") P.hb = v; else if (k == "hbytes") P.hbytes = v; else if (k == "chain") P.chain = v;
        else if (k == "nice") P.nice = v; else if (k == "lazy") P.lazy = v; else if (k == "run") P.run = v;
        else if (k == "block") P.block = v; else if (k == "unit") P.unit = v; else if (k == "group") P.group = v;
        else if (k == "minlen") P.minlen = v; else if (k == "toofar") P.toofar = v; else if (k == "good") P.good = v; else if (k == "skip2") P.skip2 = v; else if (k == "cap") P.cap = v; else if (k == "probe3") P.probe3 = v; else if (k == "lit4") P.lit4 = v;
    }
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 1;
    std::vector<uint8_t> in;
    {
        uint8_t buf[1 << 16];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) in.insert(in.end(), buf, buf + n);
    }
    fclose(f);
    size_t N = in.size();
    in.resize(N + 512, 0);
    const int maxd = 32768 - 2 * P.group;
    std::vector<uint16_t> head(1u << P.hb), prev(32768);
    std::vector<uint32_t> M(P.unit + 1);  // per position: len<<16 | dist (dist up to 32768 -> store dist-1)
    uint64_t total_bits = 0, ntok = 0, nlit = 0, searched = 0, steps = 0;
    uint64_t nblocks = 0, stored_blocks = 0, fixed_blocks = 0;
    auto load32 = [&](size_t p) { uint32_t v; memcpy(&v, &in[p], 4); return v; };
    auto hash = [&](size_t p) {
        uint32_t v = load32(p);
        if (P.hbytes == 3) v &= 0xFFFFFF;
        return (v * 2654435761u) >> (32 - P.hb);
    };
    for (size_t u0 = 0; u0 < N; u0 += P.unit) {
        size_t u1 = std::min(N, u0 + (size_t)P.unit);
        size_t h0 = u0 >= 32768 ? u0 - 32768 : 0;
        std::fill(head.begin(), head.end(), 0);
        std::fill(prev.begin(), prev.end(), 0);
        // positions are unit-relative with the history at the front: rel = p - h0 + 1 (0 means "none"... kept simple:
        // we use true u16 wrap arithmetic exactly like the device code)
        for (size_t p = h0; p < u1; ++p) {
            uint32_t h = hash(p);
            prev[p & 32767] = head[h];
            head[h] = (uint16_t)(p & 0xFFFF);
        }
        // NOTE: sequential insertion of everything first is equivalent to in-order run-ahead insertion because a
        // search only ever reads prev[] entries of positions within maxd, which a later position cannot have
        // overwritten... except that here ALL later positions are already inserted.  Re-walk safely: recompute
        // prev on the fly instead.
        std::fill(head.begin(), head.end(), 0);
        std::fill(prev.begin(), prev.end(), 0);
        size_t inserted = h0;
        auto insert_to = [&](size_t end) {
            for (; inserted < end; ++inserted) {
                uint32_t h = hash(inserted);
                prev[inserted & 32767] = head[h];
                head[h] = (uint16_t)(inserted & 0xFFFF);
            }
        };
        for (size_t g0 = u0; g0 < u1; g0 += P.group) {
            size_t g1 = std::min(u1, g0 + (size_t)P.group);
            insert_to(std::min(u1, g1 + (size_t)P.group));  // builder runs one group ahead
            int inh_len = 0, inh_dist = 0, inh_run = 0;
            for (size_t p = g0; p < g1; ++p) {
                size_t blk_end = std::min(u1, (p / P.block + 1) * (size_t)P.block);
                int maxlen = (int)std::min<size_t>(258, blk_end - p);
                int best = 0, bestd = 0;
                if ((p - g0) % P.run != 0 && inh_len - 1 >= P.minlen) {
                    best = std::min(inh_len - 1, maxlen);
                    bestd = inh_dist;
                    ++inh_run;
                } else inh_run = 0;
                if (best < P.nice && maxlen >= P.minlen) {
                    ++searched;
                    int budget = P.chain;
                    if (best >= P.good) budget >>= 2;
                    if (P.skip2 && inh_run >= 2) budget = P.skip2 >= 99 ? 0 : (P.chain >> P.skip2);
                    uint16_t c16 = prev[p & 32767];
                    uint32_t lastd = 0;
                    uint32_t cur = load32(p);
                    for (int s = 0; s < budget; ++s) {
                        uint32_t d = (uint32_t)((p - c16) & 0xFFFF);
                        if (d == 0 || d > (uint32_t)maxd || d > p - h0 || d <= lastd) break;
                        lastd = d;
                        ++steps;
                        size_t c = p - d;
                        if (load32(c) == cur || P.hbytes == 3) {
                            int l = 0;
                            int lim = P.cap ? std::min(maxlen, P.cap) : maxlen;
                            while (l < lim && in[c + l] == in[p + l]) ++l;
                            if (l > best && l >= P.minlen) {
                                best = l;
                                bestd = (int)d;
                                if (P.cap && l >= lim) {  // reached the cap: extend this one and stop
                                    while (l < maxlen && in[c + l] == in[p + l]) ++l;
                                    best = l;
                                    break;
                                }
                                if (l >= P.nice) break;
                            }
                        }
                        c16 = prev[c & 32767];
                    }
                }
                if (P.probe3 && best < 4 && maxlen >= 3) {
                    static const int pd[8] = {1, 2, 3, 4, 6, 8, 12, 16};
                    for (int k = 0; k < 8; ++k) {
                        size_t d = pd[k];
                        if (d > p - h0) break;
                        size_t c = p - d;
                        if (((load32(c) ^ load32(p)) & 0xFFFFFF) == 0) {
                            int l = 0;
                            while (l < maxlen && in[c + l] == in[p + l]) ++l;
                            if (l > best) { best = l; bestd = (int)d; }
                            break;
                        }
                    }
                }
                if (best == 3 && bestd > P.toofar) best = 0;
                if (best == 4 && P.lit4 && bestd > P.lit4) best = 0;
                inh_len = best;
                inh_dist = bestd;
                M[p - u0] = best >= std::min(P.minlen, P.probe3 ? 3 : P.minlen) ? ((uint32_t)best << 16) | (uint32_t)(bestd - 1) : 0;
            }
        }
        // parse + cost, block by block
        for (size_t b0 = u0; b0 < u1; b0 += P.block) {
            size_t b1 = std::min(u1, b0 + (size_t)P.block);
            uint32_t lf[286] = {0}, df[30] = {0};
            uint64_t extra = 0;
            size_t p = b0;
            while (p < b1) {
                uint32_t m = M[p - u0];
                int L = m >> 16;
                if (L && P.lazy && p + 1 < b1 && (int)(M[p + 1 - u0] >> 16) > L) L = 0;
                if (!L) {
                    lf[in[p]]++;
                    ++nlit;
                    ++p;
                } else {
                    int d = (m & 0xFFFF) + 1;
                    int ls = len_sym(L), ds = dist_sym(d);
                    lf[257 + ls]++;
                    df[ds]++;
                    extra += kLenExtra[ls] + kDistExtra[ds];
                    p += L;
                }
                ++ntok;
            }
            lf[256]++;
            uint8_t ll[286], dl[30];
            huff_lengths(lf, 286, 15, ll);
            huff_lengths(df, 30, 15, dl);
            uint64_t dyn = header_bits(ll, dl) + extra, fix = 3 + extra;
            for (int i = 0; i < 286; ++i) {
                dyn += (uint64_t)lf[i] * ll[i];
                fix += (uint64_t)lf[i] * (i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
            }
            for (int i = 0; i < 30; ++i) {
                dyn += (uint64_t)df[i] * dl[i];
                fix += (uint64_t)df[i] * 5;
            }
            uint64_t sto = 8 * (5 + (b1 - b0));
            uint64_t best = std::min(dyn, std::min(fix, sto));
            if (best == sto) ++stored_blocks; else if (best == fix) ++fixed_blocks;
            total_bits += ((best + 7) & ~7ull) + 40;  // pad + empty stored block (sync marker)
            ++nblocks;
        }
    }
    uint64_t bytes = total_bits / 8 + 2 + 2 + 4;
    printf("in=%zu out=%llu ratio=%.4f tokens=%llu lit=%llu searched=%.3f steps/search=%.2f blocks=%llu stored=%llu fixed=%llu\n", N,
           (unsigned long long)bytes, (double)bytes / N, (unsigned long long)ntok, (unsigned long long)nlit, (double)searched / N,
           searched ? (double)steps / searched : 0.0, (unsigned long long)nblocks, (unsigned long long)stored_blocks,
           (unsigned long long)fixed_blocks);
    return 0;
}
