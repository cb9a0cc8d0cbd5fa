// lzmodel2.cpp -- CPU model of lz77_kernel (zk_lz77.cu, round-2 design).  It follows the kernel decision for decision:
//   * exact hash chains over 4-byte hashes with 16-bit links, built one group ahead of the search;
//   * the best match M(p) of EVERY position: up to `chain` candidates that pass a one-byte test, compared over 16 bytes,
//     the longest (nearest on a tie) wins; 3-byte probes at small fixed distances where nothing was found (adaptive per
//     group); a match that fills the 16 bytes is extended exactly.  M(p) is a pure function of the data;
//   * the parse over M: greedy with one-step lazy evaluation, run speculatively from the first byte of every 16-byte
//     segment, stitched over the segment exits, with a short real parse from the true entry of every entered segment up
//     to the point where it meets the speculative one (or a cut / gap token when it does not).
// The kernel's token stream is compared with the model's token for token (tests/test_lz_model.py), and the ratio
// parameters are tuned here without a GPU.
//
// Test / design infrastructure only: nothing under zlib.kotlin_b200/ uses it.
//
//   g++ -O2 -o /tmp/lzmodel2 tools/lzmodel2.cpp && /tmp/lzmodel2 file [key=value ...]
//   g++ -O2 -shared -fPIC -DLZM2_LIB -o tools/liblzmodel2.so tools/lzmodel2.cpp     (ctypes: lzm2_tokens)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "lzcost.hpp"

namespace {

constexpr int kWindow = 32768;
constexpr int kGroup = 4096;
constexpr int kMaxDist = kWindow - 2 * kGroup;
constexpr int kHashBits = 14;
constexpr int kCap = 16;
constexpr int kMaxMatch = 258;

struct Params {
    int chain = 3, nice = 128, lazy = 1, max_lazy = 16, far3 = 256, far4 = 1024, min_len = 3, filtered = 0;
    int probe_adaptive = 1, probe_sample = 4, probe_thresh = 8;  // strips of 128 positions sampled per group / hits that switch probing on
    int seg = 16;
    int block = 32768;
    int unit = 524288;
    int lazy2 = 0;  // > 0: a match shorter than this also gives way to one at p + 2 that is at least two longer
};

struct Stats {
    uint64_t spec_searches = 0, real_searches = 0, spec_rounds = 0, real_rounds = 0, batches = 0, chain_steps = 0, compares = 0,
             probe_rounds = 0, probe_hits = 0, entered = 0, mid_entries = 0, not_met = 0, ext_bytes = 0, wasted_tokens = 0;
};

struct Match {
    int len, dist;
    bool probe_hit;
};

struct Unit {
    const uint8_t *base;  // base[q] = byte at unit-relative q (q >= h0q); readable up to data_end_q + 512
    int32_t h0q, data_end_q;
    uint32_t ulen;
    std::vector<uint16_t> prev;  // [kWindow + ulen]
    const Params *P;
    Stats *st;

    uint32_t load32(int32_t q) const {
        uint32_t v;
        memcpy(&v, base + q, 4);
        return v;
    }
    static uint32_t hash4(uint32_t v) { return (v * 0x9E3779B1u) >> (32 - kHashBits); }

    void build() {
        std::vector<uint16_t> head(1u << kHashBits, 0);
        prev.assign((size_t)kWindow + ulen + 1, 0);
        for (int32_t q = h0q; q < data_end_q; ++q) {
            const uint32_t h = hash4(load32(q));
            prev[q] = head[h];
            head[h] = (uint16_t)q;
        }
    }

    int32_t blk_end_q(int32_t q) const {
        const int32_t off = q - kWindow;
        const int32_t e = (off / P->block + 1) * P->block;
        return kWindow + std::min<int32_t>((int32_t)ulen, e);
    }

    // best match at q that is longer than best0 (0: any match of at least 3/4 bytes)
    Match search(int32_t q, int best0, bool probing) const {
        const int maxlen = std::max(0, std::min(kMaxMatch, blk_end_q(q) - q));
        const uint32_t reach = (uint32_t)std::min(kMaxDist, std::max(q - h0q, 0));
        int best = best0, bestd = 0;
        bool capped = false, hit3 = false;
        if (P->chain > 0 && maxlen >= 4 && best < maxlen && maxlen >= P->min_len) {
            const int tidx = std::max(best, 3);  // the quick test looks at one byte: the one a longer match must have
            uint32_t c16 = prev[q], lastd = 0;
            for (int step = 0; step < P->chain; ++step) {
                const uint32_t d = ((uint32_t)q - c16) & 0xFFFFu;
                if (d == 0 || d > reach || d <= lastd) break;
                lastd = d;
                const int32_t c = q - (int32_t)d;
                c16 = prev[c];
                ++st->chain_steps;
                if (base[c + tidx] == base[q + tidx]) {
                    // a candidate that differs at index `best` cannot be longer than the best so far: skipped without a
                    // compare (the kernel does the same; it does not change the result)
                    if (best >= 4 && base[c + best] != base[q + best]) continue;
                    ++st->compares;
                    const int lim = std::min(kCap, maxlen);
                    int len = 0;
                    while (len < lim && base[c + len] == base[q + len]) ++len;
                    if (len > best && len >= 4) {
                        best = len;
                        bestd = (int)d;
                        if (len >= kCap) {
                            capped = true;
                            break;
                        }
                        if (len >= P->nice || len >= maxlen) break;
                    }
                }
            }
        }
        if (probing && best < 4 && maxlen >= 3 && P->min_len <= 3) {
            static const int pd[8] = {1, 2, 3, 4, 6, 8, 12, 16};
            const uint32_t c0 = load32(q) & 0xFFFFFFu;
            int hit = 0;
            for (int k = 0; k < 8 && !hit; ++k)
                if ((load32(q - pd[k]) & 0xFFFFFFu) == c0) hit = pd[k];
            if (hit && (uint32_t)hit <= reach) {
                const int lim = std::min(kCap, maxlen);
                int len = 0;
                while (len < lim && base[q - hit + len] == base[q + len]) ++len;
                if (len >= 3 && len > best) {
                    best = len;
                    bestd = hit;
                    capped = len >= kCap;
                    hit3 = true;
                }
            }
        }
        if (capped && best < maxlen) {
            int len = best;
            while (len < maxlen && base[q - bestd + len] == base[q + len]) ++len;
            st->ext_bytes += len - best;
            best = len;
        }
        if (!bestd) best = 0;
        if (best == 3 && bestd > P->far3) best = 0;
        if (best == 4 && bestd > P->far4) best = 0;
        if (best < P->min_len || (P->filtered && best <= 5)) best = 0;
        return Match{best, best ? bestd : 0, hit3};
    }
};

struct SegRec {
    int32_t seg0 = 0, seg1 = 0, exit = 0, lastd = -1;
    uint32_t visited = 0, ismatch = 0;  // speculative parse: positions where a token starts / where it is a match
    // real prefix
    bool entered = false;
    int32_t entry = 0, meet = 0;
    uint32_t visited2 = 0, ismatch2 = 0;
    int cut_len = 0;  // > 0: the last match of the real prefix was cut to this length
    int gap_len = 0;
};

struct Out {
    std::vector<uint32_t> tokens;
    std::vector<uint32_t> blk_begin, blk_count;
};

constexpr int kStrip = 128;  // positions per search strip (the unit of the adaptive probe decision)

void run_unit(const uint8_t *base, int32_t h0q, uint32_t ulen, const Params &P, Stats &st, Out &out) {
    Unit U;
    U.base = base;
    U.h0q = h0q;
    U.ulen = ulen;
    U.data_end_q = kWindow + (int32_t)ulen;
    U.P = &P;
    U.st = &st;
    U.build();
    const int S = P.seg;
    const int nseg = kGroup / S;
    int32_t carry = kWindow;
    uint32_t blk_begin = (uint32_t)out.tokens.size();
    uint32_t prev_hits = 1u << 30;  // first group of a unit: probe everywhere
    std::vector<Match> M(kGroup + 2);
    for (int32_t g0 = kWindow; g0 < U.data_end_q; g0 += kGroup) {
        const int32_t gend = std::min(g0 + kGroup, U.data_end_q);
        // ---- the best match of every position ----
        uint32_t hits = 0;
        for (int32_t p = g0; p < gend; ++p) {
            const int strip = (p - g0) / kStrip;
            const bool probing = !P.probe_adaptive || strip < P.probe_sample || prev_hits >= (uint32_t)P.probe_thresh;
            M[p - g0] = U.search(p, 0, probing);
            ++st.spec_searches;
            if (M[p - g0].probe_hit && strip < P.probe_sample) ++hits;
            if (M[p - g0].probe_hit) ++st.probe_hits;
        }
        M[gend - g0] = Match{0, 0, false};
        M[gend - g0 + 1] = Match{0, 0, false};
        prev_hits = hits;
        // the step of the parse at p: 1 for a literal, the match length otherwise
        auto step_at = [&](int32_t p) {
            const int len = M[p - g0].len;
            if (len == 0) return 1;
            if (P.lazy && len < P.max_lazy && p + 1 < gend && M[p + 1 - g0].len > len) return 1;
            if (len < P.lazy2 && p + 2 < gend && M[p + 2 - g0].len > len + 1) return 1;
            return len;
        };
        std::vector<SegRec> rec(nseg);
        // ---- speculative parse of every segment from its first byte ----
        for (int s = 0; s < nseg; ++s) {
            SegRec &r = rec[s];
            r.seg0 = std::min(g0 + s * S, gend);
            r.seg1 = std::min(r.seg0 + S, gend);
            int32_t p = r.seg0;
            while (p < r.seg1) {
                const int stp = step_at(p);
                r.visited |= 1u << (p - r.seg0);
                if (stp > 1) {
                    r.ismatch |= 1u << (p - r.seg0);
                    r.lastd = M[p - g0].dist;
                }
                p += stp;
            }
            r.exit = p;
        }
        ++st.batches;
        // ---- stitch: which segments does the real parse enter, and where ----
        int32_t pfinal = carry;
        if (carry < gend) {
            const int first = (carry - g0) / S;
            rec[first].entered = true;
            rec[first].entry = carry;
            for (int s = first; s < nseg; ++s) {
                if (!rec[s].entered) continue;
                const int32_t x = rec[s].exit;
                if (x >= gend) {
                    pfinal = x;
                } else {
                    const int n = (x - g0) / S;
                    rec[n].entered = true;
                    rec[n].entry = x;
                }
            }
        }
        // ---- real prefix: from the entry until the parse lands on a position the speculative parse visited; it may not
        // pass the speculative exit of its segment ----
        for (int s = 0; s < nseg; ++s) {
            SegRec &r = rec[s];
            r.meet = r.seg0;
            if (!r.entered) continue;
            ++st.entered;
            if (r.entry == r.seg0) continue;
            ++st.mid_entries;
            int32_t p = r.entry;
            bool met = false;
            while (p < r.seg1) {
                if ((r.visited >> (p - r.seg0)) & 1u) {
                    met = true;
                    break;
                }
                ++st.real_searches;
                int stp = step_at(p);
                const uint32_t bit = 1u << (p - r.seg0);
                if (stp > 1 && p + stp > r.exit) {
                    stp = r.exit - p;
                    if (stp < 3) stp = 1;  // cut below a match: a literal, the parse goes on behind it
                    else r.cut_len = stp;
                }
                r.visited2 |= bit;
                if (stp > 1) r.ismatch2 |= bit;
                p += stp;
            }
            if (met) {
                r.meet = p;
            } else {
                ++st.not_met;
                r.meet = r.seg1;  // nothing of the speculative parse is used
                r.gap_len = r.exit - p;
            }
        }
        // ---- copy-out in segment order ----
        for (int s = 0; s < nseg; ++s) {
            const SegRec &r = rec[s];
            if (!r.entered) {
                st.wasted_tokens += __builtin_popcount(r.visited);
                continue;
            }
            const int last2 = r.ismatch2 ? 31 - __builtin_clz(r.ismatch2) : -1;
            for (int i = 0; i < S; ++i)
                if ((r.visited2 >> i) & 1u) {
                    const Match &m = M[r.seg0 + i - g0];
                    if ((r.ismatch2 >> i) & 1u) {
                        const int len = (i == last2 && r.cut_len) ? r.cut_len : m.len;
                        out.tokens.push_back(((uint32_t)len << 16) | (uint32_t)(m.dist - 1));
                    } else {
                        out.tokens.push_back(base[r.seg0 + i]);
                    }
                }
            const int m0 = r.meet - r.seg0;
            for (int i = m0; i < S; ++i)
                if ((r.visited >> i) & 1u) {
                    const Match &m = M[r.seg0 + i - g0];
                    if ((r.ismatch >> i) & 1u) out.tokens.push_back(((uint32_t)m.len << 16) | (uint32_t)(m.dist - 1));
                    else out.tokens.push_back(base[r.seg0 + i]);
                }
            if (r.gap_len >= 3) {
                out.tokens.push_back(((uint32_t)r.gap_len << 16) | (uint32_t)(r.lastd - 1));
            } else {
                for (int i = 0; i < r.gap_len; ++i) out.tokens.push_back(base[r.exit - r.gap_len + i]);
            }
        }
        carry = pfinal;
        const int32_t off_end = gend - kWindow;
        if (off_end == (int32_t)ulen || off_end % P.block == 0) {
            out.blk_begin.push_back(blk_begin);
            out.blk_count.push_back((uint32_t)out.tokens.size() - blk_begin);
            blk_begin = (uint32_t)out.tokens.size();
        }
    }
}

// tokens of the whole buffer (units cut at multiples of P.unit; each unit primed with the 32 KiB before it; `hist_len`
// bytes before in[0] are readable history of the first unit)
void run_all(const uint8_t *in, size_t n, uint32_t hist_len, const Params &P, Stats &st, Out &out) {
    // a padded private copy: [hist | data | zeros]
    std::vector<uint8_t> buf(64 + (size_t)hist_len + n + 1024, 0);  // 64 in front: the probes read up to 16 bytes before h0q
    if (hist_len + n) memcpy(buf.data() + 64, in - hist_len, (size_t)hist_len + n);
    const uint8_t *data = buf.data() + 64 + hist_len;
    for (size_t u0 = 0; u0 < n; u0 += (size_t)P.unit) {
        const uint32_t ulen = (uint32_t)std::min<size_t>((size_t)P.unit, n - u0);
        const int32_t h0q = u0 == 0 ? kWindow - (int32_t)hist_len : 0;
        // base[q] = data[u0 - kWindow + q]; may point before the buffer for q < h0q (never read)
        run_unit(data + (ptrdiff_t)u0 - kWindow, h0q, ulen, P, st, out);
    }
}

bool verify(const uint8_t *in, size_t n, uint32_t hist_len, const Out &out) {
    std::vector<uint8_t> rec(in - hist_len, in);
    rec.reserve(hist_len + n);
    for (uint32_t t : out.tokens) {
        const int len = (int)(t >> 16);
        if (!len) {
            rec.push_back((uint8_t)t);
        } else {
            const size_t d = (t & 0xFFFF) + 1;
            if (d > rec.size() || len < 3 || len > 258) return false;
            for (int i = 0; i < len; ++i) rec.push_back(rec[rec.size() - d]);
        }
    }
    return rec.size() == hist_len + n && memcmp(rec.data() + hist_len, in, n) == 0;
}

uint64_t cost_bits(const uint8_t *in, const Params &P, const Out &out, size_t n, uint64_t *stored_blocks) {
    uint64_t total = 0;
    size_t pos = 0;
    for (size_t b = 0; b < out.blk_begin.size(); ++b) {
        uint32_t lf[286] = {0}, df[30] = {0};
        uint64_t extra = 0;
        size_t bytes = 0;
        for (uint32_t k = 0; k < out.blk_count[b]; ++k) {
            const uint32_t t = out.tokens[out.blk_begin[b] + k];
            const int L = (int)(t >> 16);
            if (!L) {
                lf[t & 0xFF]++;
                bytes += 1;
            } else {
                const int ls = len_sym(L), ds = dist_sym((int)(t & 0xFFFF) + 1);
                lf[257 + ls]++;
                df[ds]++;
                extra += kLenExtra[ls] + kDistExtra[ds];
                bytes += L;
            }
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
        const uint64_t sto = 8 * (5 + bytes);
        const uint64_t best = std::min(dyn, std::min(fix, sto));
        if (best == sto && stored_blocks) ++*stored_blocks;
        total += ((best + 7) & ~7ull) + 40;
        pos += bytes;
    }
    (void)in;
    (void)P;
    (void)n;
    (void)pos;
    return total;
}

}  // namespace

#ifdef LZM2_LIB
// ctypes entry: tokens of in[0, n) with hist_len readable bytes in front of it.  tokens must hold n entries, blk_* must hold
// n / block + units + 1 entries.  Returns the number of tokens, or -1 when the parse does not reproduce the input.
extern "C" long long lzm2_tokens(const uint8_t *in, unsigned long long n, unsigned hist_len, const int *params, int nparams,
                                 uint32_t *tokens, uint32_t *blk_begin, uint32_t *blk_count, unsigned long long *nblocks) {
    Params P;
    int *fields[] = {&P.chain, &P.nice, &P.lazy, &P.max_lazy, &P.far3, &P.far4, &P.min_len, &P.filtered, &P.probe_adaptive,
                     &P.probe_sample, &P.probe_thresh, &P.seg, &P.block, &P.unit, &P.lazy2};
    for (int i = 0; i < nparams && i < (int)(sizeof(fields) / sizeof(fields[0])); ++i) *fields[i] = params[i];
    Stats st;
    Out out;
    run_all(in, (size_t)n, hist_len, P, st, out);
    if (!verify(in, (size_t)n, hist_len, out)) return -1;
    memcpy(tokens, out.tokens.data(), out.tokens.size() * 4);
    memcpy(blk_begin, out.blk_begin.data(), out.blk_begin.size() * 4);
    memcpy(blk_count, out.blk_count.data(), out.blk_count.size() * 4);
    *nblocks = out.blk_begin.size();
    return (long long)out.tokens.size();
}
#else
int main(int argc, char **argv) {
    if (argc < 2) return 1;
    Params P;
    for (int i = 2; i < argc; ++i) {
        char *eq = strchr(argv[i], '=');
        if (!eq) continue;
        std::string k(argv[i], eq - argv[i]);
        const int v = atoi(eq + 1);
        if (k == "chain") P.chain = v; else if (k == "nice") P.nice = v; else if (k == "lazy") P.lazy = v;
        else if (k == "max_lazy") P.max_lazy = v; else if (k == "far3") P.far3 = v; else if (k == "far4") P.far4 = v;
        else if (k == "min_len") P.min_len = v; else if (k == "probe_adaptive") P.probe_adaptive = v;
        else if (k == "probe_sample") P.probe_sample = v; else if (k == "probe_thresh") P.probe_thresh = v;
        else if (k == "seg") P.seg = v; else if (k == "lazy2") P.lazy2 = v; else if (k == "block") P.block = v; else if (k == "unit") P.unit = v;
        else { fprintf(stderr, "unknown key %s\n", k.c_str()); return 1; }
    }
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 1;
    std::vector<uint8_t> in;
    {
        static uint8_t buf[1 << 16];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) in.insert(in.end(), buf, buf + n);
    }
    fclose(f);
    Stats st;
    Out out;
    run_all(in.data(), in.size(), 0, P, st, out);
    const bool ok = verify(in.data(), in.size(), 0, out);
    uint64_t stored = 0;
    const uint64_t bits = cost_bits(in.data(), P, out, in.size(), &stored);
    const uint64_t bytes = bits / 8 + 8;
    const double N = (double)in.size();
    uint64_t nlit = 0;
    for (uint32_t t : out.tokens) nlit += (t >> 16) == 0;
    printf("valid=%d in=%zu out=%llu ratio=%.4f tokens/B=%.3f lit/B=%.3f stored_blocks=%llu\n", (int)ok, in.size(),
           (unsigned long long)bytes, bytes / N, out.tokens.size() / N, nlit / N, (unsigned long long)stored);
    printf("  searches/B spec=%.3f real=%.3f | chain steps/search=%.2f compares/search=%.2f | ext bytes/B=%.3f\n", st.spec_searches / N,
           st.real_searches / N, (double)st.chain_steps / (st.spec_searches + st.real_searches),
           (double)st.compares / (st.spec_searches + st.real_searches), st.ext_bytes / N);
    printf("  entered=%.3f of segments, mid entries=%.3f of entered, not met=%.4f of mid, wasted spec tokens/B=%.3f\n",
           (double)st.entered / (st.batches * (double)(kGroup / P.seg)), (double)st.mid_entries / st.entered, (double)st.not_met / std::max<uint64_t>(1, st.mid_entries),
           st.wasted_tokens / N);
    return ok ? 0 : 2;
}
#endif
