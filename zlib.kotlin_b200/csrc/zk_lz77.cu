// zk_lz77.cu -- K1: LZ77 match finding + lazy parse, one CTA per LZ unit.
//
// Replaces the tokeniser of DeflateStream.compressZlibDynamic (deflate/DeflateStream.kt:573-790:
// hash3 DS:588-595, insertTripletAt DS:127-144, chain walk DS:668-692, advanceMatch DS:165-184)
// with a design that maps to an SM instead of a scalar loop.  It is not a port: the reference's
// look-ahead/window arrays, its pre-seeded chains and its stale-window compare do not exist here.
//
// Round-2 design: the parse drives the search.  Round 1 searched every position and parsed the match array
// afterwards (17.5 warp instructions per input byte); here a lane owns a short segment of the input and walks it
// the way a scalar compressor would -- search where the parse lands, one look-ahead search for the lazy rule,
// jump over the match -- so that positions inside a match cost nothing.
//
// Layout in shared memory (one CTA per SM, ~200 KB):
//   ring  u8 [65536]   input bytes at (q & 0xFFFF); q = unit-relative position, history first
//   prev  u16[32768]   hash-chain links at (q & 32767): previous position with the same 4-byte hash
//   head  u16[16384]   most recent position (q & 0xFFFF) per 14-bit hash
//   ph/pi [2][4096]    per position of a group to be built: its hash, and which lower lane of its batch shares it
//   mt/mt2 u32[6][256] match tokens of every segment: speculative parse / real prefix
//
// Software pipeline over groups of 4096 positions, one __syncthreads per step.  In step i
//   - one elected thread stages group i+4 into the ring with a TMA bulk copy (mbarrier-tracked),
//   - the prep warps hash group i+2 and resolve equal hashes inside every batch of 32 positions (match.any),
//   - the builder warp inserts group i+1 into the hash chains: table updates only, eight batches in flight,
//   - the eight search warps turn group i into tokens.  Lane L owns segment L (16 bytes):
//       speculative parse  from the first byte of the segment, in warp-lockstep rounds of one search per lane;
//       stitch             the real parse enters the segment the carry points into and, from an entered segment, the
//                          one its speculative exit lies in: pointer doubling with shuffles, two team barriers;
//       real prefix        an entered lane whose entry is not its first byte parses from the entry until it lands on a
//                          position the speculative parse visited (usually at once: every literal is such a position);
//       copy-out           prefix sum of the token counts, tokens leave in order.
// A search only follows links of candidates within 32768 - 2*4096 bytes, which the builder (one group ahead) can
// not have overwritten: results do not depend on warp timing.  tools/lzmodel2.cpp is a CPU model of exactly this
// algorithm; tests compare the two token for token.
//
// HBM traffic per input byte: 1 B read (+ 32 KiB halo per unit) and 4 B x tokens/byte written.
// The symbol histograms are built from the token stream by hist_kernel (zk_huffman.cu).
#include "zk_common.cuh"

namespace zk {

namespace {

constexpr int kHashBits = 14;  // 16384 buckets of 16-bit positions (q & 0xFFFF: the ring holds 64 KiB, links are 16-bit too)
constexpr int kSeg = 16;                       // bytes per lane segment
constexpr int kSegs = kGroup / kSeg;           // segments (= search lanes) per group
constexpr int kSearchWarps = kSegs / 32;
constexpr int kPrepWarps = 3;
constexpr int kBuilderWarp = kSearchWarps;     // warp index of the builder
constexpr int kFirstPrepWarp = kSearchWarps + 1;
constexpr int kWarpsLz = kSearchWarps + 1 + kPrepWarps;
constexpr int kThreadsLz = 32 * kWarpsLz;
constexpr int kMaxM = (kSeg + 2) / 3;          // matches that can start inside one segment
constexpr int kHistGroups = kWindow / kGroup;
constexpr int kMaxDist = kWindow - 2 * kGroup;
constexpr int kStages = 4;                     // ring staging barriers
constexpr int kCap = 16;  // candidates are compared over 16 bytes; the nearest one that fills them is extended
// The short-distance probes (3-byte matches of binary records and runs) find nothing in text: a warp probes in its
// first kProbeRounds rounds of a group, and in all of them once it has had kProbeThresh hits in this group or had them
// in the group before.  The decision is a function of the data of the warp's own segments only.
constexpr int kProbeRounds = 2;
constexpr int kProbeThresh = 2;
constexpr uint32_t kFull = 0xFFFFFFFFu;

static_assert(kSeg <= 32 && kSegs % 32 == 0, "one visited-position mask per segment, whole warps per group");

struct LzSmem {
    uint32_t ring[65536 / 4 + 8];  // + 32 mirrored bytes so a 20-byte read never wraps
    uint16_t prev[32768];
    uint16_t head[1 << kHashBits];  // most recent position (q & 0xFFFF) per hash
    uint16_t ph[2][kGroup];         // hash of every position of a group to be built
    uint8_t pi[2][kGroup];          // 0x80 last of its hash in its batch, 0x40 | lane of the nearest lower lane with the same hash
    uint32_t mt[kMaxM][kSegs];      // match tokens of the speculative parse of segment L, in order
    uint32_t mt2[kMaxM][kSegs];     // match tokens of the real prefix of segment L
    uint32_t jmark[kSegs];          // stitch: segments of L's warp the parse visits from L on
    int16_t jnext[kSegs];           // stitch: first segment after L's warp the parse from L leads to
    uint16_t pstart[kSegs];         // entry position (relative to the group) of an entered segment
    unsigned long long bar[kStages];
    int wcount[kSearchWarps];
    int32_t pfinal;
    int32_t carry_entry[2];         // first unparsed q, by group parity
    uint32_t carry_tok_off[2];      // tokens emitted so far in this unit
    uint32_t carry_blk_begin[2];    // carry_tok_off at the start of the current block
};

struct LzArgs {
    const uint8_t *in;      // shard input
    uint64_t n;             // shard length
    const uint8_t *hist;    // bytes preceding the shard (hist_len of them), or nullptr
    uint32_t hist_len;
    uint32_t unit_bytes;    // multiple of block_bytes
    uint32_t block_bytes;   // power of two, multiple of kGroup
    uint32_t block_shift;   // log2(block_bytes)
    uint32_t num_units;
    LevelParams lp;
    uint32_t *tokens;       // [n] dense per unit, unit u starts at u * unit_bytes
    uint32_t *block_tok_begin;  // [num_blocks] index into tokens
    uint32_t *block_tok_count;  // [num_blocks]
    unsigned long long *prof;   // optional [num_units][32] cycle counters (nullptr = off)
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 4 bytes at byte offset q of the ring (any alignment, wraps at 64 KiB)
__device__ __forceinline__ uint32_t ring_load32(const uint32_t *ring, uint32_t q) {
    const uint32_t w = (q >> 2) & 0x3FFFu;
    const uint32_t lo = ring[w], hi = ring[w + 1];  // the mirror behind the ring covers the wrap
    return __funnelshift_r(lo, hi, (q & 3u) << 3);
}
__device__ __forceinline__ uint32_t ring_load8(const uint32_t *ring, uint32_t q) {
    return reinterpret_cast<const uint8_t *>(ring)[q & 0xFFFFu];
}
__device__ __forceinline__ uint32_t hash4(uint32_t four_bytes) { return (four_bytes * 0x9E3779B1u) >> (32 - kHashBits); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// TMA bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

struct UnitGeom {
    uint64_t unit_start;  // shard offset of the unit's first byte
    uint32_t ulen;        // bytes in the unit
    int32_t h0q;          // first valid q (history start)
    int32_t data_end_q;   // kWindow + ulen
    int first_group;      // group holding h0q
    int last_group;       // last data group
};

// Stage group t of the unit into the ring.  Source bytes may come from the shard itself or, for the
// history of unit 0, from the separate history buffer.  Returns true when the TMA path was used.
__device__ __forceinline__ bool stage_group(LzSmem &s, const LzArgs &a, const UnitGeom &g, int t) {
    const int32_t q0 = t * kGroup;
    // valid source range of this group in q coordinates
    int32_t lo = max(q0, g.h0q);
    // bytes past the unit end are only ever hashed, never matched: stage what the input still has
    const int64_t avail_q = (int64_t)kWindow + (int64_t)tmin<uint64_t>(a.n - g.unit_start, (uint64_t)g.ulen + kGroup);
    int32_t hi = (int32_t)tmin<int64_t>((int64_t)q0 + kGroup, avail_q);
    if (hi <= lo) return false;
    const uint8_t *src;
    if (g.unit_start == 0 && q0 < kWindow) {
        src = a.hist + (int64_t)a.hist_len - kWindow;  // q = 0 maps to hist_len - 32768 (may be "before" hist)
    } else {
        src = a.in + g.unit_start - kWindow;
    }
    uint8_t *ring8 = reinterpret_cast<uint8_t *>(s.ring);
    const bool tma = lo == q0 && ((hi - lo) & 15) == 0 && ((reinterpret_cast<uintptr_t>(src + lo)) & 15) == 0;
    if (tma) {
        if (threadIdx.x == kFirstPrepWarp * 32) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&s.bar[t & (kStages - 1)], (uint32_t)(hi - lo));
            tma_load(ring8 + (q0 & 0xFFFF), src + lo, (uint32_t)(hi - lo), &s.bar[t & (kStages - 1)]);
        }
    } else {
        for (int32_t q = lo + (int32_t)threadIdx.x; q < hi; q += kThreadsLz) ring8[q & 0xFFFF] = src[q];
    }
    return tma;
}

// ---- preparation of group t, shared by the prep warps -----------------------------------------------------
// For every position its hash, which lower lane of its 32-position batch has the same hash, and whether it is the last
// of its hash in the batch (an exact match.any).  A pure function of the batch, so it runs in parallel and the serial
// builder is left with table updates only.
__device__ __forceinline__ void prep_group(LzSmem &s, int t, int pw, int lane) {
    uint16_t *hh = s.ph[t & 1];
    uint8_t *pp = s.pi[t & 1];
    const uint32_t q0 = (uint32_t)t * kGroup + lane;
    constexpr int kBatches = kGroup / 32;
#pragma unroll 2
    for (int b = pw; b < kBatches; b += kPrepWarps) {
        const uint32_t hv = hash4(ring_load32(s.ring, q0 + b * 32));
        const uint32_t eq = __match_any_sync(kFull, hv);
        const uint32_t below = eq & ((1u << lane) - 1u);
        uint32_t info = below ? (0x40u | (uint32_t)(31 - __clz(below))) : 0u;
        if (((eq >> lane) >> 1) == 0u) info |= 0x80u;
        hh[b * 32 + lane] = (uint16_t)hv;
        pp[b * 32 + lane] = (uint8_t)info;
    }
}

// ---- builder: one warp inserts group t into the hash chains ----------------------------------
// Exact, deterministic chains at LSU issue rate: per batch of 32 positions read head[h] (the link for lanes that have
// no lower lane with their hash in the batch), then the last lane of every hash stores its position.  The shared
// accesses of one warp are handled in issue order, so the next batch's read sees this batch's store; the loaded links
// are consumed only after eight batches have been issued.  No votes, no branches, no read-backs.
__device__ __forceinline__ void build_group(LzSmem &s, int t, int lane) {
    const uint32_t q0 = (uint32_t)t * kGroup + lane;
    volatile uint16_t *vhead = s.head;
    const uint16_t *hh = s.ph[t & 1] + lane;
    const uint8_t *pp = s.pi[t & 1] + lane;
#pragma unroll 1
    for (int b = 0; b < kGroup / 32; b += 8) {
        uint32_t h[8], old[8], info[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            h[u] = hh[(b + u) * 32];
            info[u] = pp[(b + u) * 32];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            old[u] = vhead[h[u]];
            if (info[u] & 0x80u) vhead[h[u]] = (uint16_t)(q0 + (b + u) * 32);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t q = q0 + (b + u) * 32;
            const uint32_t link = (info[u] & 0x40u) ? (q - lane) + (info[u] & 31u) : old[u];
            s.prev[q & 32767u] = (uint16_t)link;
        }
    }
    __syncwarp();
}

// ---- search: the best match at one position ----------------------------------------------------------------
// 16 bytes at q as four words
struct Bytes16 {
    uint32_t w[4];
};
__device__ __forceinline__ Bytes16 ring_load128(const uint32_t *ring, uint32_t q) {
    const uint32_t i = (q >> 2) & 0x3FFFu, sh = (q & 3u) << 3;
    const uint32_t *r5 = ring + i;  // words i..i+4; the mirror behind the ring covers the wrap
    const uint32_t a0 = r5[0], a1 = r5[1], a2 = r5[2], a3 = r5[3], a4 = r5[4];
    Bytes16 r;
    r.w[0] = __funnelshift_r(a0, a1, sh);
    r.w[1] = __funnelshift_r(a1, a2, sh);
    r.w[2] = __funnelshift_r(a2, a3, sh);
    r.w[3] = __funnelshift_r(a3, a4, sh);
    return r;
}
// number of equal leading bytes of two 16-byte strings (0..16)
__device__ __forceinline__ int common_prefix16(const Bytes16 &a, const Bytes16 &b) {
    // branch-free: pick the first differing word with selects, then one 32-bit find-first-set
    const uint32_t x0 = a.w[0] ^ b.w[0], x1 = a.w[1] ^ b.w[1], x2 = a.w[2] ^ b.w[2], x3 = a.w[3] ^ b.w[3];
    const uint32_t w01 = x0 ? x0 : x1, w23 = x2 ? x2 : x3;
    const int b01 = x0 ? 0 : 4, b23 = x2 ? 8 : 12;
    const bool lowhalf = (x0 | x1) != 0u;
    const uint32_t w = lowhalf ? w01 : w23;
    const int base = lowhalf ? b01 : b23;
    return w ? base + ((__ffs((int)w) - 1) >> 3) : 16;
}

struct SearchCtx {
    LevelParams lp;
    int32_t h0q;
    int32_t ulen;
    uint32_t block_shift;
};

struct Found {
    int len, dist;
    bool hit3;
};

// The best match at q that is longer than best0 (0: any match).  Called by all lanes of a warp; `act` = this lane wants
// an answer.  Every loop is uniform across the warp (lanes drop out by predicate).
__device__ __forceinline__ Found search(const LzSmem &s, const SearchCtx &cx, bool act, int32_t q, int best0, bool probing) {
    const LevelParams &lp = cx.lp;
    // matches never cross the end of their deflate block
    const int32_t off = q - kWindow;
    const int32_t blk_end = min(cx.ulen, ((off >> cx.block_shift) + 1) << cx.block_shift);
    const int maxlen = act ? min(kMaxMatch, blk_end - off) : 0;
    const uint32_t reach = (uint32_t)min(kMaxDist, q - cx.h0q);
    const Bytes16 cur = ring_load128(s.ring, (uint32_t)q);
    int best = best0, bestd = 0;
    bool capped = false;  // the best match filled the 16-byte compare and may be longer
    // ---- hash-chain candidates.  The walk only applies a one-byte test (the byte a longer match must have); the
    // candidates that pass are compared over 16 bytes afterwards, nearest first.
    {
        bool walk = lp.chain > 0 && maxlen >= 4 && best < maxlen && maxlen >= lp.min_len;
        const uint32_t tidx = (uint32_t)max(best0, 3);
        const uint32_t need = ring_load8(s.ring, (uint32_t)q + tidx);
        uint32_t c16 = s.prev[q & 32767];
        uint32_t lastd = 0;
        for (int base = 0; base < lp.chain; base += 4) {
            if (!__any_sync(kFull, walk)) break;
            unsigned long long passed = 0;  // up to four 16-bit distances, nearest in the low bits
            int npass = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t d = ((uint32_t)q - c16) & 0xFFFFu;
                walk = walk && base + k < lp.chain && d - 1u < reach && d > lastd;
                if (walk) {
                    lastd = d;
                    const uint32_t c = (uint32_t)q - d;
                    c16 = s.prev[c & 32767u];
                    if (ring_load8(s.ring, c + tidx) == need) {
                        passed |= (unsigned long long)d << (16 * npass);
                        ++npass;
                    }
                }
            }
            while (__any_sync(kFull, passed != 0ull)) {
                if (passed != 0ull) {
                    const uint32_t d = (uint32_t)passed & 0xFFFFu;
                    passed >>= 16;
                    const int len = min(common_prefix16(cur, ring_load128(s.ring, (uint32_t)q - d)), maxlen);
                    if (len > best && len >= 4) {
                        best = len;
                        bestd = (int)d;
                        if (len >= kCap) {
                            capped = true;
                            walk = false;
                            passed = 0ull;
                        } else if (len >= lp.nice || len >= maxlen) {
                            walk = false;
                            passed = 0ull;
                        }
                    }
                }
            }
        }
    }
    // ---- short matches at small fixed distances (3-byte repeats of binary records, byte runs) -------------
    bool hit3 = false;
    {
        const bool want = probing && best < 4 && maxlen >= 3 && lp.min_len <= 3;
        if (__any_sync(kFull, want)) {
            // the 3 bytes at q against the 3 bytes at q-1, -2, -3, -4, -6, -8, -12, -16: all come from the 16 bytes before q
            const Bytes16 bk = ring_load128(s.ring, (uint32_t)(q - 16));
            const uint32_t c0 = cur.w[0] & 0xFFFFFFu;
            int hit = 0;
            if ((bk.w[0] & 0xFFFFFFu) == c0) hit = 16;
            if ((bk.w[1] & 0xFFFFFFu) == c0) hit = 12;
            if ((bk.w[2] & 0xFFFFFFu) == c0) hit = 8;
            if ((__funnelshift_r(bk.w[2], bk.w[3], 16) & 0xFFFFFFu) == c0) hit = 6;
            if ((bk.w[3] & 0xFFFFFFu) == c0) hit = 4;
            if ((__funnelshift_r(bk.w[3], cur.w[0], 8) & 0xFFFFFFu) == c0) hit = 3;
            if ((__funnelshift_r(bk.w[3], cur.w[0], 16) & 0xFFFFFFu) == c0) hit = 2;
            if ((__funnelshift_r(bk.w[3], cur.w[0], 24) & 0xFFFFFFu) == c0) hit = 1;
            if (want && hit && (uint32_t)hit <= reach) {
                const int len = min(common_prefix16(cur, ring_load128(s.ring, (uint32_t)(q - hit))), maxlen);
                if (len >= 3 && len > best) {
                    best = len;
                    bestd = hit;
                    capped = len >= kCap;
                    hit3 = true;
                }
            }
        }
    }
    // ---- a match that filled the 16-byte compare is extended, 16 bytes per trip -------------------------------
    {
        bool ext = capped && best < maxlen;
        int len = kCap;
        while (__any_sync(kFull, ext)) {
            if (ext) {
                const int n = common_prefix16(ring_load128(s.ring, (uint32_t)(q + len)), ring_load128(s.ring, (uint32_t)(q + len - bestd)));
                len += n;
                if (n < 16 || len >= maxlen) {
                    ext = false;
                    best = min(len, maxlen);
                }
            }
        }
    }
    if (bestd == 0) best = 0;  // nothing longer than best0
    if (best == 3 && bestd > lp.far3) best = 0;
    if (best == 4 && bestd > lp.far4) best = 0;
    if (best < lp.min_len || (lp.filtered && best <= 5)) best = 0;
    Found f;
    f.len = best;
    f.dist = bestd;
    f.hit3 = hit3;
    return f;
}

// ---- the parse of one lane ----------------------------------------------------------------------------------
struct ParseState {
    int32_t p;       // current position
    int phase;       // 1: a match at p is pending (curlen, curdist) and the search looks at p + 1
    int curlen, curdist;
    uint32_t visited, ismatch;  // positions of the segment where a token starts / where that token is a match
    int nm;          // matches emitted
    int32_t lastd;   // distance of the last emitted match
};

// One round: one search, then the decision (literal, match, or wait for the look-ahead).  kReal: the parse of an entered
// segment from its true entry; it may not pass `stop`, the speculative exit of the segment.
template <bool kReal>
__device__ __forceinline__ bool lane_round(LzSmem &s, const SearchCtx &cx, bool act, ParseState &st, uint32_t (*mt)[kSegs], int L,
                                           int32_t seg0, int32_t limit, int32_t gend, int32_t stop, bool probing) {
    const LevelParams &lp = cx.lp;
    const Found R = search(s, cx, act, st.p + st.phase, st.phase ? st.curlen : 0, probing);
    if (!act) return false;
    bool emit = false;  // emit the match (curlen, curdist) at st.p
    if (st.phase == 0) {
        if (R.len == 0) {
            st.visited |= 1u << (st.p - seg0);
            st.p += 1;
        } else {
            st.curlen = R.len;
            st.curdist = R.dist;
            if (lp.lazy && R.len < lp.max_lazy && st.p + 1 < gend) st.phase = 1;
            else emit = true;
        }
    } else {
        if (R.len > st.curlen) {
            // the byte at p becomes a literal, the longer match at p + 1 is the pending one
            st.visited |= 1u << (st.p - seg0);
            st.p += 1;
            st.curlen = R.len;
            st.curdist = R.dist;
            if (st.p >= limit) st.phase = 0;  // the pending match belongs to the next segment
            else if (!(R.len < lp.max_lazy && st.p + 1 < gend)) emit = true;
        } else {
            emit = true;
        }
    }
    if (emit) {
        int len = st.curlen;
        st.phase = 0;
        if (kReal && st.p + len > stop) {
            len = stop - st.p;
            if (len < 3) {  // cut below a match: a literal, the parse goes on behind it
                st.visited |= 1u << (st.p - seg0);
                st.p += 1;
                return R.hit3;
            }
        }
        st.visited |= 1u << (st.p - seg0);
        st.ismatch |= 1u << (st.p - seg0);
        mt[st.nm][L] = ((uint32_t)len << 16) | (uint32_t)(st.curdist - 1);
        st.nm += 1;
        st.lastd = st.curdist;
        st.p += len;
    }
    return R.hit3;
}

__device__ __forceinline__ void team_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kSearchWarps * 32) : "memory"); }

// The search team turns group t into tokens.  L = team lane = segment index.  prev_hits: probe hits of this warp in the
// group before (warp-uniform, carried in a register).
__device__ __forceinline__ void parse_group(LzSmem &s, const LzArgs &a, const UnitGeom &g, const SearchCtx &cx, int t, int L,
                                            uint32_t unit_index, uint32_t &prev_hits, long long (&pt)[8]) {
    long long tc = clock64();
#define ZK_PT(i)                         \
    {                                    \
        const long long now = clock64(); \
        pt[i] += now - tc;               \
        tc = now;                        \
    }
    const int lane = L & 31, sw = L >> 5;
    const int32_t g0 = t * kGroup;
    const int32_t gend = min(g0 + kGroup, g.data_end_q);
    const int32_t seg0 = min(g0 + L * kSeg, gend), seg1 = min(seg0 + kSeg, gend);
    const bool adaptive = cx.lp.probe_adaptive != 0;

    // ---- speculative parse of the segment from its first byte ----
    ParseState sp;
    sp.p = seg0;
    sp.phase = 0;
    sp.curlen = sp.curdist = 0;
    sp.visited = sp.ismatch = 0u;
    sp.nm = 0;
    sp.lastd = -1;
    uint32_t hits = 0;
    {
        int round = 0;
        for (;;) {
            const bool act = sp.p < seg1;
            if (!__any_sync(kFull, act)) break;
            const bool probing = !adaptive || round < kProbeRounds || hits >= kProbeThresh || prev_hits >= kProbeThresh;
            const bool hit3 = lane_round<false>(s, cx, act, sp, s.mt, L, seg0, seg1, gend, 0, probing);
            hits += __popc(__ballot_sync(kFull, hit3));
            ++round;
        }
        pt[6] += round;
    }
    prev_hits = hits;
    const bool probing_b = !adaptive || hits >= kProbeThresh;
    const int32_t my_exit = sp.p;  // >= seg1; an empty segment (seg0 == gend) exits at gend
    ZK_PT(0)

    // ---- stitch.  The real parse enters the segment that holds the carry, and from an entered segment l the one its
    // speculative exit lies in (the real parse of a segment is made to end at the speculative exit, below): which
    // segments are entered is reachability along those links.  Level 1, inside each warp with shuffles (pointer
    // doubling, 5 rounds): for every segment of the warp the set of the warp's segments visited from it (a 32-bit
    // mask) and the first segment after the warp it leads to.  Level 2, after one barrier: every lane walks the hops
    // from warp to warp.
    const int n0 = my_exit >= gend ? kSegs : (my_exit - g0) / kSeg;
    {
        const int wbase = sw * 32;
        int jl = n0;
        uint32_t vm = 1u << lane;
#pragma unroll
        for (int round = 0; round < 5; ++round) {
            const bool local = jl < wbase + 32;
            const uint32_t vn = __shfl_sync(kFull, vm, jl & 31);
            const int jn = __shfl_sync(kFull, jl, jl & 31);
            if (local) {
                vm |= vn;
                jl = jn;
            }
        }
        s.jnext[L] = (int16_t)jl;
        s.jmark[L] = vm;
    }
    const int par = t & 1;
    const int32_t x0 = s.carry_entry[par];
    const int first = x0 >= gend ? kSegs : (x0 - g0) / kSeg;
    if (L == 0) s.pfinal = x0;  // stays when nothing is entered
    team_barrier();
    bool entered = false;
    {
        int cur = first;
#pragma unroll 1
        while (cur < kSegs) {
            if ((cur >> 5) == sw) entered = (s.jmark[cur] >> lane) & 1u;
            cur = s.jnext[cur];
        }
        if (L == first) s.pstart[L] = (uint16_t)(x0 - g0);
        if (entered) {
            if (n0 < kSegs) s.pstart[n0] = (uint16_t)(my_exit - g0);
            else s.pfinal = my_exit;
        }
    }
    team_barrier();
    ZK_PT(1)

    // ---- real prefix: from the entry until the parse lands on a position the speculative parse visited ----
    ParseState rp;
    rp.p = entered ? g0 + (int32_t)s.pstart[L] : seg0;
    rp.phase = 0;
    rp.curlen = rp.curdist = 0;
    rp.visited = rp.ismatch = 0u;
    rp.nm = 0;
    rp.lastd = -1;
    bool met = true;  // an entry at the first byte of the segment meets at once
    {
        bool act = entered && rp.p > seg0;
        if (act) met = false;
        int round = 0;
        for (;;) {
            if (act) {
                if (rp.p >= seg1) {
                    act = false;
                } else if ((sp.visited >> (rp.p - seg0)) & 1u) {
                    met = true;
                    act = false;
                }
            }
            if (!__any_sync(kFull, act)) break;
            lane_round<true>(s, cx, act, rp, s.mt2, L, seg0, seg1, gend, my_exit, probing_b);
            ++round;
        }
        pt[7] += round;
    }
    // tokens of this lane: the real prefix [entry, meet), the speculative tokens from meet on, and when the two never
    // met the rest up to the speculative exit: that gap is the tail of the speculative parse's last match
    const int m = met ? rp.p - seg0 : kSeg;           // first speculative position used
    const int gap = met ? 0 : my_exit - rp.p;         // bytes between the real parse's end and the speculative exit
    const uint32_t spec_mask = m >= 32 ? 0u : sp.visited & ~((1u << m) - 1u);
    int cnt = 0;
    if (entered) cnt = __popc(rp.visited) + __popc(spec_mask) + (gap >= 3 ? 1 : gap);
    // dense offsets across the team
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == 31) s.wcount[sw] = incl;
    ZK_PT(2)
    team_barrier();
    ZK_PT(3)
    int before = 0, total = 0;
#pragma unroll
    for (int k = 0; k < kSearchWarps; ++k) {
        const int c = s.wcount[k];
        if (k < sw) before += c;
        total += c;
    }
    const uint32_t tok_off = s.carry_tok_off[par];
    if (entered) {
        uint32_t *dst = a.tokens + (uint64_t)unit_index * a.unit_bytes + tok_off + (uint32_t)(before + incl - cnt);
        {
            uint32_t v = rp.visited;
            int k = 0;
            while (v) {
                const int i = __ffs((int)v) - 1;
                v &= v - 1u;
                uint32_t tok;
                if ((rp.ismatch >> i) & 1u) tok = s.mt2[k++][L];
                else tok = ring_load8(s.ring, (uint32_t)(seg0 + i));
                *dst++ = tok;
            }
        }
        {
            uint32_t v = spec_mask;
            int k = m >= 32 ? 0 : __popc(sp.ismatch & ((1u << m) - 1u));
            while (v) {
                const int i = __ffs((int)v) - 1;
                v &= v - 1u;
                uint32_t tok;
                if ((sp.ismatch >> i) & 1u) tok = s.mt[k++][L];
                else tok = ring_load8(s.ring, (uint32_t)(seg0 + i));
                *dst++ = tok;
            }
        }
        if (gap >= 3) {
            *dst++ = ((uint32_t)gap << 16) | (uint32_t)(sp.lastd - 1);
        } else {
            for (int i = 0; i < gap; ++i) *dst++ = ring_load8(s.ring, (uint32_t)(my_exit - gap + i));
        }
    }
    // close the deflate block when this group ends it, then hand the carry to the next group (other parity: no lane
    // of this group reads what is written here)
    if (L == 0) {
        const int32_t off_end = gend - kWindow;
        const bool closes = off_end == (int32_t)g.ulen || (off_end & (int32_t)(a.block_bytes - 1)) == 0;
        const uint32_t new_off = tok_off + (uint32_t)total;
        uint32_t blk_begin = s.carry_blk_begin[par];
        if (closes) {
            const uint32_t blk = (uint32_t)((g.unit_start + (uint64_t)(off_end - 1)) >> a.block_shift);
            a.block_tok_begin[blk] = (uint32_t)((uint64_t)unit_index * a.unit_bytes) + blk_begin;
            a.block_tok_count[blk] = new_off - blk_begin;
            blk_begin = new_off;
        }
        s.carry_entry[par ^ 1] = s.pfinal;
        s.carry_tok_off[par ^ 1] = new_off;
        s.carry_blk_begin[par ^ 1] = blk_begin;
    }
    ZK_PT(4)
#undef ZK_PT
}

__global__ void __launch_bounds__(kThreadsLz, 1) lz77_kernel(const LzArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    LzSmem &s = *reinterpret_cast<LzSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool is_search = warp < kSearchWarps;
    const bool is_builder = warp == kBuilderWarp;
    const bool is_prep = warp >= kFirstPrepWarp;

    for (uint32_t unit = blockIdx.x; unit < a.num_units; unit += gridDim.x) {
        UnitGeom g;
        g.unit_start = (uint64_t)unit * a.unit_bytes;
        g.ulen = (uint32_t)tmin<uint64_t>(a.unit_bytes, a.n - g.unit_start);
        g.h0q = unit == 0 ? (int32_t)(kWindow - a.hist_len) : 0;
        g.data_end_q = kWindow + (int32_t)g.ulen;
        g.first_group = g.h0q / kGroup;
        g.last_group = (g.data_end_q - 1) / kGroup;
        SearchCtx cx;
        cx.lp = a.lp;
        cx.h0q = g.h0q;
        cx.ulen = (int32_t)g.ulen;
        cx.block_shift = a.block_shift;

        // reset tables (deterministic start) and barriers
        {
            uint4 *z = reinterpret_cast<uint4 *>(s.prev);
            const int nvec = (int)((sizeof(s.prev) + sizeof(s.head)) / 16);
            for (int i = threadIdx.x; i < nvec; i += kThreadsLz) z[i] = make_uint4(0, 0, 0, 0);
            if (threadIdx.x == 0) {
                for (int b = 0; b < kStages; ++b) mbar_init(&s.bar[b], 1);
                s.carry_entry[0] = s.carry_entry[1] = kWindow;
                s.carry_tok_off[0] = s.carry_tok_off[1] = 0;
                s.carry_blk_begin[0] = s.carry_blk_begin[1] = 0;
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
        }
        __syncthreads();
        // stage uses per barrier slot, to derive the parity of the next wait
        uint32_t stage_uses[kStages] = {0, 0, 0, 0};
        // prologue: groups first_group, first_group+1 (the loop stages i+4 = first_group+2 onwards)
        bool staged_tma[2];
        for (int k = 0; k < 2; ++k) staged_tma[k] = stage_group(s, a, g, g.first_group + k);
        for (int k = 0; k < 2; ++k) {
            if (staged_tma[k]) {
                const int slot = (g.first_group + k) & (kStages - 1);
                uint32_t spins = 0;
                while (!mbar_try_wait(&s.bar[slot], stage_uses[slot] & 1u)) {
                    if (++spins > (1u << 26)) break;  // never hang the device on a lost transaction
                }
                stage_uses[slot]++;
            }
        }
        __syncthreads();
        if (threadIdx.x < 8) s.ring[65536 / 4 + threadIdx.x] = s.ring[threadIdx.x];
        __syncthreads();

        // step i: stage(i+4) | prepare(i+2) | build(i+1) | search + parse(i)
        uint32_t prev_hits = 1u << 30;  // the first group of a unit has no group before it: probe everywhere
        long long t_role = 0;
        long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const long long t_unit0 = clock64();
        for (int i = g.first_group - 2; i <= g.last_group; ++i) {
            const bool tma_next = (i + 4 <= g.last_group + 1) ? stage_group(s, a, g, i + 4) : false;
            const long long t0 = clock64();
            if (is_prep && i + 2 >= g.first_group && i + 2 <= g.last_group) prep_group(s, i + 2, warp - kFirstPrepWarp, lane);
            if (is_builder && i + 1 <= g.last_group && i + 1 >= g.first_group) build_group(s, i + 1, lane);
            if (is_search && i >= kHistGroups && i <= g.last_group) parse_group(s, a, g, cx, i, warp * 32 + lane, unit, prev_hits, pt);
            t_role += clock64() - t0;
            if (tma_next) {
                const int slot = (i + 4) & (kStages - 1);
                uint32_t spins = 0;
                while (!mbar_try_wait(&s.bar[slot], stage_uses[slot] & 1u)) {
                    if (++spins > (1u << 26)) break;
                }
                stage_uses[slot]++;
            }
            if ((((i + 4) * kGroup) & 0xFFFF) == 0) {  // the group staged in this step sits at ring offset 0
                __syncthreads();
                if (threadIdx.x < 8) s.ring[65536 / 4 + threadIdx.x] = s.ring[threadIdx.x];
            }
            __syncthreads();
        }
        if (a.prof && lane == 0 && (warp == 0 || warp == kBuilderWarp || warp == kFirstPrepWarp)) {
            // per unit: [0] total; [1] builder busy; [2] prep busy; [3] search team busy; [4..8] its phases (speculative
            // parse, stitch, real prefix + count, wait, copy-out); [9] speculative rounds, [10] real rounds (warp 0)
            unsigned long long *o = a.prof + (size_t)unit * 32;
            if (warp == kBuilderWarp) {
                o[0] = (unsigned long long)(clock64() - t_unit0);
                o[1] = (unsigned long long)t_role;
            } else if (warp == kFirstPrepWarp) {
                o[2] = (unsigned long long)t_role;
            } else {
                o[3] = (unsigned long long)t_role;
                for (int k = 0; k < 5; ++k) o[4 + k] = (unsigned long long)pt[k];
                o[9] = (unsigned long long)pt[6];
                o[10] = (unsigned long long)pt[7];
            }
        }
    }
}

}  // namespace

LevelParams level_params(int level, int strategy) {
    // chain / nice / lazy per level.  The reference only varies the chain length (8/16/32/64 for level
    // <=2/<=4/<=6/else, DS:603-609) and never parses lazily (DS:610); these are tuned with tools/lzmodel2.cpp so that
    // level 6 lands just below libz level-6 sizes (4-byte hashes in 16384 buckets make a chain step worth several of the
    // reference's 3-byte ones).  max_lazy is zlib's: a match at least that long is taken without looking at the next
    // position.
    LevelParams p;
    p.far3 = 256;
    p.far4 = 1024;
    p.min_len = 3;
    p.filtered = strategy == ZK_FILTERED;
    p.probe_adaptive = level <= 7;
    p.nice = 128;
    p.lazy = 1;
    p.max_lazy = 16;
    switch (level) {
        case 1: p.chain = 1; p.lazy = 0; p.nice = 32; break;
        case 2: p.chain = 2; p.lazy = 0; p.nice = 32; break;
        case 3: p.chain = 2; p.nice = 64; p.max_lazy = 8; break;
        case 4: p.chain = 2; break;
        case 5: p.chain = 3; p.nice = 64; break;
        case 6: p.chain = 3; break;
        case 7: p.chain = 6; p.max_lazy = 32; break;
        case 8: p.chain = 16; p.nice = 258; p.max_lazy = 128; break;
        default: p.chain = 32; p.nice = 258; p.max_lazy = 258; break;
    }
    if (level < 1) p.chain = 1;
    if (strategy == ZK_HUFFMAN_ONLY) {
        p.chain = 0;
        p.min_len = kMaxMatch + 1;
    }
    return p;
}

size_t lz77_smem_bytes() { return sizeof(LzSmem); }

int launch_lz77(zk_ctx *c, const uint8_t *d_in, uint64_t n, const uint8_t *d_hist, uint32_t hist_len,
                uint32_t unit_bytes, uint32_t block_bytes, const LevelParams &lp, uint32_t *tokens,
                uint32_t *block_tok_begin, uint32_t *block_tok_count, unsigned long long *prof) {
    LzArgs a;
    a.prof = prof;
    a.in = d_in;
    a.n = n;
    a.hist = d_hist;
    a.hist_len = d_hist ? hist_len : 0;
    a.unit_bytes = unit_bytes;
    a.block_bytes = block_bytes;
    a.block_shift = 0;
    while ((1u << a.block_shift) < block_bytes) ++a.block_shift;
    a.num_units = (uint32_t)((n + unit_bytes - 1) / unit_bytes);
    a.lp = lp;
    a.tokens = tokens;
    a.block_tok_begin = block_tok_begin;
    a.block_tok_count = block_tok_count;
    // per-device attribute: set on every launch (cheap) so multi-device processes stay correct
    ZK_CUDA(cudaFuncSetAttribute(lz77_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LzSmem)));
    const uint32_t grid = a.num_units;
    lz77_kernel<<<grid, kThreadsLz, sizeof(LzSmem), c->stream>>>(a);
    c->launches += 1;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}

}  // namespace zk
