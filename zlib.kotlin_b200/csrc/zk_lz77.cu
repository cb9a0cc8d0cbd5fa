// zk_lz77.cu -- K1: LZ77 match finding + lazy parse, one CTA per LZ unit.
//
// Replaces the tokeniser of DeflateStream.compressZlibDynamic (deflate/DeflateStream.kt:573-790:
// hash3 DS:588-595, insertTripletAt DS:127-144, chain walk DS:668-692, advanceMatch DS:165-184)
// with a design that maps to an SM instead of a scalar loop.  It is not a port: the reference's
// look-ahead/window arrays, its pre-seeded chains and its stale-window compare do not exist here.
//
// Layout in shared memory (one CTA per SM, ~224 KB):
//   ring  u8 [65536]   input bytes at (q & 0xFFFF); q = unit-relative position, history first
//   prev  u16[32768]   hash-chain links at (q & 32767): previous position with the same 4-byte hash
//   head  u16[16384]   most recent position (q & 0xFFFF) per 14-bit hash
//   match u32[2][4608] per-position best match of the group in flight / the group being parsed (rows of 32 skewed)
//   prep  u8 [2][4096] per position of a group to be built: lower lane of its batch with the same hash
//   stage u32[34][128], pre u32[4][128]  tokens of the group being parsed, one column per parser lane
//
// Software pipeline over groups of 4096 positions, one __syncthreads per step.  In step i
//   - one elected thread stages group i+4 into the ring with a TMA bulk copy (mbarrier-tracked),
//   - whichever warp takes a strip prepares the same strip of group i+2 (match.any over the hashes of each batch),
//   - warp 31 inserts group i+1 into the hash chains: table updates only, 32 positions per instruction sequence,
//   - warps 24..27 (one per SM sub-partition) parse group i-1: a register-only speculative pass per 32-byte lane
//     segment, a two-level stitch with shuffles, a short real pass up to the point where it meets the speculative
//     one, then the tokens leave densely with 16-byte stores,
//   - every warp (the role warps join when done) steals strips of group i: 40 of 64 positions (a lane walks 2
//     consecutive positions, the second inherits (len-1, dist) from the first), then 48 of 32; the chain loop is
//     uniform across the warp: each candidate is compared over 16 bytes with no inner loop, the nearest one that
//     fills them is extended afterwards, 3-byte matches come from probes at small fixed distances (adaptive).
// A search only follows links of candidates within 32768 - 2*4096 bytes, which the builder (one
// group ahead) can not have overwritten: results do not depend on warp timing.
//
// HBM traffic per input byte: 1 B read (+ 32 KiB halo per unit) and 4 B x tokens/byte written.
// The symbol histograms are built from the token stream by hist_kernel (zk_huffman.cu).
#include "zk_common.cuh"

#ifndef ZK_USE_MATCH_ANY
#define ZK_USE_MATCH_ANY 1
#endif

namespace zk {

namespace {

constexpr int kThreadsLz = 1024;
constexpr int kHashBits = 14;  // 16384 buckets of 16-bit positions (q & 0xFFFF: the ring holds 64 KiB, links are 16-bit too)
constexpr int kRun = 2;                      // consecutive positions per lane (strip = 64 positions)
constexpr int kStrip = 32 * kRun;            // positions per work-stealing strip
// Work-stealing strips of a group: the first kBigStrips cover 64 positions each (two per lane: the second inherits from the
// first), the rest of the group is cut into strips of 32 positions.  A step ends one strip after the last role warp has
// joined the search, so the strips handed out last are the short ones.
#ifndef ZK_BIG_STRIPS
#define ZK_BIG_STRIPS 40
#endif
constexpr int kBigStrips = ZK_BIG_STRIPS;
constexpr int kSmallStrips = (kGroup - kBigStrips * kStrip) / 32;
constexpr int kStripsPerGroup = kBigStrips + kSmallStrips;
constexpr int kParserWarps = 4;
constexpr int kSeg = kGroup / (32 * kParserWarps);  // parser: positions per lane segment
constexpr int kPreRows = 4;                         // parser: real-parse tokens a lane may emit before convergence
// The sequential roles: one parser warp per SM sub-partition (warps 24..27) and the builder on warp 31.  With the heavy
// roles of the first version all five had to share one sub-partition whose other warps waited; with today's lean roles
// that one issue port had become the bottleneck.
constexpr int kHistGroups = kWindow / kGroup;
constexpr int kMaxDist = kWindow - 2 * kGroup;
constexpr int kStages = 4;                   // ring staging barriers
// The short-distance probes (3-byte matches of binary records and runs) cost a fifth of the search and find nothing in
// text: every group probes in its first kProbeSample strips only, and in all of them when the sample of the group
// before it had at least kProbeThreshold hits.  Strips are fixed ranges of positions, so the decision does not depend
// on which warp happens to search them.
constexpr int kProbeSample = 8;
constexpr int kProbeThreshold = 8;

struct LzSmem {
    uint32_t ring[65536 / 4 + 8];  // + 32 mirrored bytes so a 20-byte read never wraps
    uint16_t prev[32768];
    uint16_t head[1 << kHashBits];  // most recent position (q & 0xFFFF) per hash
    // best match per position of a group, skewed: position i sits at mvi(i) = i + 4 * (i / 32), so that the parser lanes
    // (one 32-position segment each) read their rows with conflict-free 16-byte loads; + slack: one word past the group
    uint32_t match[2][kGroup + kGroup / 8 + 8];
    uint8_t prep[2][kGroup];  // per position of a group to be built: 0x80 last of its hash in its batch, 0x40 | lane of
                              // the nearest lower lane with the same hash
    uint32_t stage[(kSeg + 2) * kParserWarps * 32];  // tokens of the group being parsed: [k][parser lane]
    uint32_t pre[kPreRows * kParserWarps * 32];      // tokens of the real parse before it meets the speculative one
    int32_t jnext[kParserWarps * 32];                // stitch: segment the parse of segment l exits into (pointer jumping)
    int32_t jmark[kParserWarps * 32];                // stitch: segment l is entered by the real parse
    unsigned long long bar[kStages];
    int work[2];
    int probe_hits[4];  // short-distance probe hits in the sampled strips of group t, at [t & 3]
    int32_t pstart[kParserWarps * 32];
    int32_t pstop[kParserWarps * 32];
    int32_t pfinal;
    int32_t plast[kParserWarps * 32];  // distance-1 of the speculative parse's last match (closes a gap)
    int pcount[kParserWarps];
    int32_t carry_entry;       // first unparsed q
    uint32_t carry_tok_off;    // tokens emitted so far in this unit
    uint32_t carry_blk_begin;  // carry_tok_off at the start of the current block
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
    unsigned long long *prof;   // optional [num_units][8] cycle counters (nullptr = off)
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 4 bytes at byte offset q of the ring (any alignment, wraps at 64 KiB)
__device__ __forceinline__ uint32_t ring_load32(const uint32_t *ring, uint32_t q) {
    const uint32_t w = (q >> 2) & 0x3FFFu;
    const uint32_t lo = ring[w], hi = ring[(w + 1) & 0x3FFFu];
    return __funnelshift_r(lo, hi, (q & 3u) << 3);
}
__device__ __forceinline__ uint32_t ring_load8(const uint32_t *ring, uint32_t q) {
    return (ring[(q >> 2) & 0x3FFFu] >> ((q & 3u) << 3)) & 0xFFu;
}
__device__ __forceinline__ int mvi(int i) { return i + ((i >> 5) << 2); }
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
__device__ __forceinline__ bool stage_group(LzSmem &s, const LzArgs &a, const UnitGeom &g, int t, bool issue) {
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
        if (issue && threadIdx.x == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&s.bar[t & (kStages - 1)], (uint32_t)(hi - lo));
            tma_load(ring8 + (q0 & 0xFFFF), src + lo, (uint32_t)(hi - lo), &s.bar[t & (kStages - 1)]);
        }
    } else if (issue) {
        for (int32_t q = lo + (int32_t)threadIdx.x; q < hi; q += kThreadsLz) ring8[q & 0xFFFF] = src[q];
    }
    return tma;
}

// ---- preparation of a strip of group t, done by whichever warp searches the matching strip ------------------
// For every position: which lower lane of its 32-position batch has the same hash, and whether it is the last of its
// hash in the batch (an exact match.any; cheap here because 26 warps share its latency, unlike in the serial builder
// where it was first tried).  It is a pure function of the batch, so it runs in parallel on the searcher warps and the
// serial builder is left with table updates only.
__device__ __forceinline__ void prep_strip(LzSmem &s, int t, int pos0, int batches, int lane) {
    uint8_t *pp = s.prep[t & 1] + pos0;
    const uint32_t q0 = (uint32_t)t * kGroup + pos0 + lane;
#pragma unroll 1
    for (int u = 0; u < batches; ++u) {
        const uint32_t hv = hash4(ring_load32(s.ring, q0 + u * 32));
#if ZK_USE_MATCH_ANY
        const uint32_t eq = __match_any_sync(0xFFFFFFFFu, hv);
#else
        uint32_t eq = 0xFFFFFFFFu;
#pragma unroll
        for (int bit = 0; bit < kHashBits; ++bit) {
            const bool one = (hv >> bit) & 1u;
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, one);
            eq &= one ? bal : ~bal;
        }
#endif
        const uint32_t below = eq & ((1u << lane) - 1u);
        uint32_t info = below ? (0x40u | (uint32_t)(31 - __clz(below))) : 0u;
        if (((eq >> lane) >> 1) == 0u) info |= 0x80u;
        pp[u * 32 + lane] = (uint8_t)info;
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
    const uint8_t *pp = s.prep[t & 1] + lane;
#pragma unroll 1
    for (int b = 0; b < kGroup / 32; b += 8) {
        uint32_t h[8], old[8], info[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            h[u] = hash4(ring_load32(s.ring, q0 + (b + u) * 32));
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

// ---- searcher: best match of kRun consecutive positions per lane ----------------------------------
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

constexpr int kCap = 16;  // candidates are compared over 16 bytes; the nearest one that reaches 16 is extended

__device__ __forceinline__ void search_strip(LzSmem &s, const LevelParams &lp, const UnitGeom &g, int t, int strip,
                                             int pos0, int run, int lane, uint32_t block_shift, bool probe_all) {
    const bool probing = probe_all || strip < kProbeSample;
    uint32_t *mv = s.match[t & 1];
    const int32_t base = t * kGroup + pos0 + lane * run;
    int inh_len = 0, inh_dist = 0, inh_run = 0;
#pragma unroll 1
    for (int r = 0; r < run; ++r) {
        const int32_t p = base + r;
        const bool valid = p < g.data_end_q;
        // matches never cross the end of their deflate block
        const int32_t off = p - kWindow;
        const int32_t blk_end = min((int32_t)g.ulen, ((off >> block_shift) + 1) << block_shift);
        const int maxlen = valid ? min(kMaxMatch, blk_end - off) : 0;
        int best = 0, bestd = 0;
        if (r > 0 && inh_len - 1 >= 3) {
            best = min(inh_len - 1, maxlen);
            bestd = inh_dist;
            ++inh_run;
        } else {
            inh_run = 0;
        }
        int budget = lp.chain;
        if (best >= lp.good) budget >>= 2;
        if (lp.skip_deep && inh_run >= 2) budget = 0;
        // an inherited match that already fills the 16-byte compare cannot be beaten by it
        if (best >= lp.nice || best >= kCap || best >= maxlen || maxlen < lp.min_len) budget = 0;
        const uint32_t reach = (uint32_t)min(kMaxDist, max(p - g.h0q, 0));
        const Bytes16 cur = ring_load128(s.ring, (uint32_t)p);
        // ---- hash-chain candidates: the loop is uniform across the warp, lanes drop out by predicate --------
        bool capped = false;  // the best match filled the 16-byte compare and may be longer
        {
            uint32_t c16 = s.prev[p & 32767];
            uint32_t lastd = 0;
            bool open = budget > 0;
            for (int step = 0; step < lp.chain; ++step) {
                const uint32_t d = ((uint32_t)p - c16) & 0xFFFFu;
                open = open && step < budget && d != 0u && d <= reach && d > lastd;
                if (!__any_sync(0xFFFFFFFFu, open)) break;
                if (open) {
                    lastd = d;
                    const uint32_t c = (uint32_t)p - d;
                    c16 = s.prev[c & 32767u];
                    // only a candidate that also matches the byte at index `best` can be longer than the best so far
                    int len = 0;
                    const uint32_t cw = best < 8 ? (best < 4 ? cur.w[0] : cur.w[1]) : (best < 12 ? cur.w[2] : cur.w[3]);
                    const bool may_win = ring_load8(s.ring, c + (uint32_t)best) == ((cw >> ((best & 3) << 3)) & 0xFFu);
                    if (may_win) len = min(common_prefix16(cur, ring_load128(s.ring, c)), maxlen);
                    if (len > best && len >= 4) {
                        best = len;
                        bestd = (int)d;
                        capped = len >= kCap;
                        if (capped || len >= lp.nice) open = false;
                    }
                }
            }
        }
        // ---- short matches at small fixed distances (3-byte repeats of binary records, byte runs) -------------
        {
            const bool want = probing && valid && best < 4 && maxlen >= 3 && lp.min_len <= 3;
            if (__any_sync(0xFFFFFFFFu, want)) {
                bool hit3 = false;
                // the 3 bytes at p against the 3 bytes at p-1, -2, -3, -4, -6, -8, -12, -16: all come from the 16
                // bytes before p
                const Bytes16 bk = ring_load128(s.ring, (uint32_t)(p - 16));
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
                    const int len = min(common_prefix16(cur, ring_load128(s.ring, (uint32_t)(p - hit))), maxlen);
                    if (len >= 3) {
                        best = len;
                        bestd = hit;
                        capped = len >= kCap;
                        hit3 = true;
                    }
                }
                if (strip < kProbeSample) {
                    const int hits = __popc(__ballot_sync(0xFFFFFFFFu, hit3));
                    if (lane == 0 && hits) atomicAdd(&s.probe_hits[t & 3], hits);
                }
            }
        }
        // ---- extension of matches that filled the 16-byte compare: cooperative, one round per distinct distance -----
        // For a distance d the end of every match of the strip is the first x with byte[x] != byte[x-d]; the lanes of
        // the warp compare 3 x 32 words (384 bytes from the strip's first position + 16) and ballot the mismatching
        // words, then each lane with that distance finds the first mismatch at or after its own position + 16.
        {
            bool ext = capped && best < maxlen;
            uint32_t pending = __ballot_sync(0xFFFFFFFFu, ext);
            const uint32_t scan0 = ((uint32_t)(t * kGroup + pos0 + r) + kCap) & ~3u;  // word aligned
            while (pending) {
                const int leader = __ffs(pending) - 1;
                const uint32_t d = (uint32_t)__shfl_sync(0xFFFFFFFFu, bestd, leader);
                const bool mine = ext && (uint32_t)bestd == d;
                uint32_t diff[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const uint32_t x = scan0 + (uint32_t)(k * 32 + lane) * 4u;
                    diff[k] = __ballot_sync(0xFFFFFFFFu, ring_load32(s.ring, x) != ring_load32(s.ring, x - d));
                }
                if (mine) {
                    // first mismatching word at or after my position + 16
                    const uint32_t from = (uint32_t)p + kCap;            // first byte not yet compared
                    const uint32_t w0 = (from - scan0) >> 2;             // its word index in the bitmap (0..~20)
                    int len = maxlen;
                    uint32_t wfound = 96;
#pragma unroll
                    for (int k = 2; k >= 0; --k) {
                        uint32_t bits = diff[k];
                        const int lo = (int)w0 - k * 32;
                        if (lo > 0) bits = lo >= 32 ? 0u : bits & ~((1u << lo) - 1u);
                        if (bits) wfound = (uint32_t)(k * 32 + __ffs(bits) - 1);
                    }
                    if (wfound < 96) {
                        // the exact byte: compare from max(from, word start)
                        const uint32_t wx = scan0 + wfound * 4u;
                        const uint32_t start = wx > from ? wx : from;
                        const uint32_t xr = ring_load32(s.ring, start) ^ ring_load32(s.ring, start - d);
                        // bytes [start, start+4): the first differing one (there is one inside the word; if `from`
                        // cut the word and the difference lies before it, the next word decides -- handled by the
                        // conservative fallback below)
                        if (xr) len = (int)(start - (uint32_t)p) + ((__ffs((int)xr) - 1) >> 3);
                        else len = (int)(start - (uint32_t)p) + 4;
                        if (!xr) {
                            // rare: the mismatch of that word was before `from`; continue byte-exact with a short loop
                            int l2 = len;
                            while (l2 < maxlen) {
                                const uint32_t y = ring_load32(s.ring, (uint32_t)p + l2) ^ ring_load32(s.ring, (uint32_t)p + l2 - d);
                                if (y) {
                                    l2 += (__ffs((int)y) - 1) >> 3;
                                    break;
                                }
                                l2 += 4;
                            }
                            len = l2;
                        }
                    }
                    best = min(len, maxlen);
                    ext = false;
                }
                pending &= ~__ballot_sync(0xFFFFFFFFu, mine);
            }
        }
        if (best == 3 && bestd > lp.far3) best = 0;
        if (best == 4 && bestd > lp.far4) best = 0;
        inh_len = best;
        inh_dist = bestd;
        uint32_t word = 0;
        if (best >= lp.min_len && !(lp.filtered && best <= 5)) word = ((uint32_t)best << 16) | (uint32_t)(bestd - 1);
        mv[mvi(p - t * kGroup)] = word;
    }
}

// ---- parser: four warps turn the per-position matches of group t into tokens -------------------------------
__device__ __forceinline__ int tok_len(uint32_t w) { return (int)((w >> 16) & 0x1FFu); }

// Decide the token at p: returns the step (1 for a literal, len for a match).
__device__ __forceinline__ int decide(const uint32_t *mv, int32_t g0, int32_t gend, int32_t p, int lazy) {
    // both words are loaded unconditionally (one slot of slack behind the group keeps p + 1 in bounds): no
    // dependent second load, no branch
    const int len = tok_len(mv[mvi(p - g0)]);
    const int next = tok_len(mv[mvi(p + 1 - g0)]);
    const bool defer = lazy && p + 1 < gend && next > len;
    return (len == 0 || defer) ? 1 : len;
}

__device__ __forceinline__ void parser_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kParserWarps * 32) : "memory"); }

// L = parser lane 0..127.  The carry (first unparsed position, tokens emitted, block start) lives in shared
// memory and is advanced by parser lane 0 after the last barrier of the group.
__device__ __forceinline__ void parse_group(LzSmem &s, const LzArgs &a, const UnitGeom &g, int t, int L,
                                            uint32_t unit_index, long long (&pt)[6]) {
    long long tc = clock64();
#define ZK_PT(i)                        \
    {                                   \
        const long long now = clock64(); \
        pt[i] += now - tc;              \
        tc = now;                       \
    }
    uint32_t *mv = s.match[t & 1];
    const int lane = L & 31, pw = L >> 5;
    const int32_t g0 = t * kGroup;
    const int32_t gend = min(g0 + kGroup, g.data_end_q);
    const int lazy = a.lp.lazy;
    const int32_t seg0 = min(g0 + L * kSeg, gend), seg1 = min(seg0 + kSeg, gend);

    constexpr int kLanes = kParserWarps * 32;
    static_assert(kSeg == 32, "the visited-position mask of a segment is one 32-bit word");
    uint32_t *col = s.stage + L;
    // pass A: speculative parse of this lane's segment from its first position.  The tokens go to the lane's staging
    // column, the positions it visits into a bit mask, the exit (and the distance of the match that makes it) to the
    // stitch.
    // No dependent chain through shared memory: the 33 match words of the segment are loaded with independent
    // 16-byte loads, the step of every position (the lazy rule) is packed 8 bits per position into registers, and
    // "which positions does the parse visit" is 32 unrolled register operations.  Tokens are then written for the
    // visited positions (independent, predicated).
    uint32_t visited = 0;
    int cnt_spec = 0;
    int32_t my_exit = seg0;
    {
        const int seg_len = seg1 - seg0;
        int32_t lastd = -1;
        if (seg_len > 0) {
            const uint32_t *mp = mv + mvi(seg0 - g0);  // the segment's 32 words are contiguous
            uint32_t st[8];
            uint4 cur = *reinterpret_cast<const uint4 *>(mp);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint4 nxt;
                if (c < 7) nxt = *reinterpret_cast<const uint4 *>(mp + 4 * c + 4);
                else nxt = make_uint4(mv[mvi(seg0 - g0 + 32)], 0u, 0u, 0u);
                const uint32_t w[5] = {cur.x, cur.y, cur.z, cur.w, nxt.x};
                uint32_t packed = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int len = tok_len(w[j]), nextlen = tok_len(w[j + 1]);
                    const bool defer = lazy && seg0 + 4 * c + j + 1 < gend && nextlen > len;
                    const uint32_t step = (len == 0 || defer) ? 1u : (uint32_t)min(len, 64);
                    packed |= step << (8 * j);
                }
                st[c] = packed;
                cur = nxt;
            }
            visited = 1u;
#pragma unroll
            for (int i = 0; i < 31; ++i) {
                const uint32_t tgt = (uint32_t)i + ((st[i >> 2] >> (8 * (i & 3))) & 0xFFu);
                if (((visited >> i) & 1u) && tgt < 32u) visited |= 1u << tgt;
            }
            if (seg_len < 32) visited &= (1u << seg_len) - 1u;
            const int lastpos = 31 - __clz(visited);
            {
                const int step = decide(mv, g0, gend, seg0 + lastpos, lazy);
                lastd = step > 1 ? (int32_t)(mp[lastpos] & 0xFFFFu) : -1;
                my_exit = seg0 + lastpos + step;
            }
            cnt_spec = __popc(visited);
            // the segment's 32 bytes (aligned in the ring) for the literal tokens
            const uint4 *rp = reinterpret_cast<const uint4 *>(s.ring + (((uint32_t)seg0 >> 2) & 0x3FFFu));
            const uint4 r0 = rp[0], r1 = rp[1];
            const uint32_t rb[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
            // eight positions at a time: the match words first (two 16-byte loads), then the stores -- a store to the
            // staging column could alias the match array as far as the compiler knows, so one load per store would wait
            // for shared memory 32 times in a row
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint4 wa = *reinterpret_cast<const uint4 *>(mp + 8 * c), wb = *reinterpret_cast<const uint4 *>(mp + 8 * c + 4);
                const uint32_t w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int i = 8 * c + j;
                    if ((visited >> i) & 1u) {
                        const bool is_match = ((st[i >> 2] >> (8 * (i & 3))) & 0xFFu) > 1u;
                        const uint32_t tok = is_match ? (w8[j] & 0x01FFFFFFu) : ((rb[i >> 2] >> (8 * (i & 3))) & 0xFFu);
                        col[__popc(visited & ((1u << i) - 1u)) * kLanes] = tok;
                    }
                }
            }
        }
        s.plast[L] = lastd;
    }
    // stitch.  The real parse enters the segment that holds the carry, and from an entered segment l the one its
    // speculative exit lies in (the two parses have met by then, or the gap is closed below): which segments are entered
    // is reachability along those links.  Level 1, inside each parser warp with shuffles (pointer doubling, 5 rounds): for
    // every segment of the warp the set of the warp's segments visited from it (a 32-bit mask) and the first segment
    // after the warp it leads to.  Level 2, after one barrier: every lane walks the at most four hops from warp to warp.
    const int n0 = my_exit >= gend ? kLanes : (my_exit - g0) / kSeg;
    {
        const int wbase = pw * 32;
        int jl = n0;
        uint32_t vm = 1u << lane;
#pragma unroll
        for (int round = 0; round < 5; ++round) {
            const bool local = jl < wbase + 32;
            const uint32_t vn = __shfl_sync(0xFFFFFFFFu, vm, jl & 31);
            const int jn = __shfl_sync(0xFFFFFFFFu, jl, jl & 31);
            if (local) {
                vm |= vn;
                jl = jn;
            }
        }
        s.jnext[L] = jl;
        s.jmark[L] = (int32_t)vm;
    }
    ZK_PT(0)
    const int32_t x0 = s.carry_entry;
    const int first = x0 >= gend ? kLanes : (x0 - g0) / kSeg;
    if (L == 0) s.pfinal = x0;  // stays when nothing is entered
    parser_barrier();
    ZK_PT(1)
    {
        bool entered = false;
        int cur = first;
#pragma unroll 1
        while (cur < kLanes) {
            if ((cur >> 5) == pw) entered = (((uint32_t)s.jmark[cur]) >> lane) & 1u;
            cur = s.jnext[cur];
        }
        if (L == first) s.pstart[L] = x0;
        if (entered) {
            if (n0 < kLanes) s.pstart[n0] = my_exit;
            else s.pfinal = my_exit;
        }
        parser_barrier();
        s.pstop[L] = entered ? my_exit : 0;
        if (!entered) s.pstart[L] = 0;
    }
    ZK_PT(2)
    const int32_t my_start = s.pstart[L], my_stop = s.pstop[L];
    // pass B: the lane's tokens are those of the real parse of [my_start, my_stop).  It usually meets the speculative
    // parse within a few tokens (at a position that one visited): those few go to the prefix rows, the rest is already in
    // the column.  Otherwise the range is parsed again into the column; when the real parse leaves the segment before
    // my_stop, the rest is the tail of the speculative parse's last match: one match at that distance, or up to two
    // literals, closes the gap.
    int pre_cnt = 0, skip = 0, cnt = 0;
    if (my_start < my_stop) {
        int32_t p = my_start;
        bool met = false;
#pragma unroll 1
        for (int it = 0; it <= kPreRows; ++it) {
            if (p >= seg1) break;
            if ((visited >> (p - seg0)) & 1u) {
                met = true;
                break;
            }
            if (it == kPreRows) break;
            const uint32_t w = mv[mvi(p - g0)];
            const int step = decide(mv, g0, gend, p, lazy);
            s.pre[pre_cnt * kLanes + L] = step > 1 ? (((uint32_t)step << 16) | (w & 0xFFFFu)) : ring_load8(s.ring, (uint32_t)p);
            ++pre_cnt;
            p += step;
        }
        if (met) {
            skip = __popc(visited & ((1u << (p - seg0)) - 1u));
            cnt = pre_cnt + cnt_spec - skip;
        } else {
            pre_cnt = 0;
            skip = 0;
            p = my_start;
            while (p < my_stop) {
                uint32_t tok;
                int step;
                if (p >= seg1) {
                    const int gap = my_stop - p;
                    const int32_t d1 = s.plast[L];
                    if (gap >= 3 && d1 >= 0) {
                        tok = ((uint32_t)gap << 16) | (uint32_t)d1;
                        step = gap;
                    } else {
                        tok = ring_load8(s.ring, (uint32_t)p);
                        step = 1;
                    }
                } else {
                    const uint32_t w = mv[mvi(p - g0)];
                    step = decide(mv, g0, gend, p, lazy);
                    if (p + step > my_stop) step = my_stop - p;  // overshoot: cut (a match of < 3 becomes literals)
                    if (step >= 3) {
                        tok = ((uint32_t)step << 16) | (w & 0xFFFFu);
                    } else {
                        tok = ring_load8(s.ring, (uint32_t)p);
                        step = 1;
                    }
                }
                col[cnt * kLanes] = tok;
                p += step;
                ++cnt;
            }
        }
    }
    // dense offsets across the 128 parser lanes
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == 31) s.pcount[pw] = incl;
    ZK_PT(3)
    parser_barrier();
    ZK_PT(4)
    int before = 0, total = 0;
#pragma unroll
    for (int k = 0; k < kParserWarps; ++k) {
        const int c = s.pcount[k];
        if (k < pw) before += c;
        total += c;
    }
    const uint32_t tok_off = s.carry_tok_off;
    // copy-out: independent loads from the staging column, dense stores (adjacent lanes write adjacent runs)
    {
        uint32_t *dst = a.tokens + (uint64_t)unit_index * a.unit_bytes + tok_off + (uint32_t)(before + incl - cnt);
        for (int k = 0; k < pre_cnt; ++k) dst[k] = s.pre[k * kLanes + L];
        const uint32_t *src = col + skip * kLanes;
        dst += pre_cnt;
        const int rest = cnt - pre_cnt;
        int k = 0;
        // a lane's tokens are one contiguous run in global memory: scalar up to a 16-byte boundary, then four at a time
        for (; k < rest && ((reinterpret_cast<uintptr_t>(dst + k)) & 15u) != 0; ++k) dst[k] = src[k * kLanes];
        for (; k + 4 <= rest; k += 4) {
            const uint32_t t0 = src[k * kLanes], t1 = src[(k + 1) * kLanes], t2 = src[(k + 2) * kLanes], t3 = src[(k + 3) * kLanes];
            *reinterpret_cast<uint4 *>(dst + k) = make_uint4(t0, t1, t2, t3);
        }
        for (; k < rest; ++k) dst[k] = src[k * kLanes];
    }
    // close the deflate block when this group ends it, then advance the carry (one lane; the next group's parse
    // starts after the CTA-wide barrier of the step)
    const int32_t off_end = gend - kWindow;
    const bool closes = off_end == (int32_t)g.ulen || (off_end & (int32_t)(a.block_bytes - 1)) == 0;
    parser_barrier();  // every parser lane has read the carry
    if (L == 0) {
        const uint32_t new_off = tok_off + (uint32_t)total;
        if (closes) {
            const uint32_t blk = (uint32_t)((g.unit_start + (uint64_t)(off_end - 1)) >> a.block_shift);
            a.block_tok_begin[blk] = (uint32_t)((uint64_t)unit_index * a.unit_bytes) + s.carry_blk_begin;
            a.block_tok_count[blk] = new_off - s.carry_blk_begin;
            s.carry_blk_begin = new_off;
        }
        s.carry_entry = s.pfinal;
        s.carry_tok_off = new_off;
    }
    ZK_PT(5)
#undef ZK_PT
}

__global__ void __launch_bounds__(kThreadsLz, 1) lz77_kernel(const LzArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    LzSmem &s = *reinterpret_cast<LzSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // one parser warp per sub-partition (warps 24..27), the builder on warp 31; nobody waits
    const bool is_builder = warp == 31;
    const bool is_parser = warp >= 24 && warp < 28;
    const int parser_index = warp - 24;

    for (uint32_t unit = blockIdx.x; unit < a.num_units; unit += gridDim.x) {
        UnitGeom g;
        g.unit_start = (uint64_t)unit * a.unit_bytes;
        g.ulen = (uint32_t)tmin<uint64_t>(a.unit_bytes, a.n - g.unit_start);
        g.h0q = unit == 0 ? (int32_t)(kWindow - a.hist_len) : 0;
        g.data_end_q = kWindow + (int32_t)g.ulen;
        g.first_group = g.h0q / kGroup;
        g.last_group = (g.data_end_q - 1) / kGroup;

        // reset tables (deterministic start) and barriers
        {
            uint4 *z = reinterpret_cast<uint4 *>(s.prev);
            const int nvec = (int)((sizeof(s.prev) + sizeof(s.head)) / 16);
            for (int i = threadIdx.x; i < nvec; i += kThreadsLz) z[i] = make_uint4(0, 0, 0, 0);
            if (threadIdx.x == 0) {
                for (int b = 0; b < kStages; ++b) mbar_init(&s.bar[b], 1);
                s.work[0] = s.work[1] = 0;
                s.probe_hits[0] = s.probe_hits[1] = s.probe_hits[2] = s.probe_hits[3] = 0;
                s.carry_entry = kWindow;
                s.carry_tok_off = 0;
                s.carry_blk_begin = 0;
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
        }
        __syncthreads();
        // stage uses per barrier slot, to derive the parity of the next wait
        uint32_t stage_uses[kStages] = {0, 0, 0, 0};
        // prologue: groups first_group, first_group+1 (the loop stages i+4 = first_group+2 onwards)
        bool staged_tma[2];
        for (int k = 0; k < 2; ++k) staged_tma[k] = stage_group(s, a, g, g.first_group + k, true);
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

        // step i: stage(i+4) | prepare(i+2) | build(i+1) | search(i) | parse(i-1)
        long long t_role = 0, t_search = 0;
        long long pt[6] = {0, 0, 0, 0, 0, 0};
        const long long t_unit0 = clock64();
        for (int i = g.first_group - 2; i <= g.last_group + 1; ++i) {
            const bool tma_next = (i + 4 <= g.last_group + 1) ? stage_group(s, a, g, i + 4, true) : false;
            if (threadIdx.x == 0) {
                s.work[(i + 1) & 1] = 0;
                s.probe_hits[(i + 1) & 3] = 0;
            }
            // (the first group of a unit has no sample before it: it probes everywhere)
            const bool probe_all = !a.lp.probe_adaptive || i <= kHistGroups || s.probe_hits[(i - 1) & 3] >= kProbeThreshold;

            const long long t0 = clock64();
            if (is_builder && i + 1 <= g.last_group && i + 1 >= g.first_group) build_group(s, i + 1, lane);
            if (is_parser && i - 1 >= kHistGroups && i - 1 <= g.last_group)
                parse_group(s, a, g, i - 1, parser_index * 32 + lane, unit, pt);
            // the other warps of the role sub-partition stay off its issue port until the builder and the parsers have
            // finished this step: a named barrier (roles arrive, guests wait), no polling
            const long long t1 = clock64();
            const bool searching = i >= kHistGroups && i <= g.last_group;
            const bool preparing = i + 2 >= g.first_group && i + 2 <= g.last_group;
            if (searching || preparing) {
                for (;;) {
                    int strip = 0;
                    if (lane == 0) strip = atomicAdd(&s.work[i & 1], 1);
                    strip = __shfl_sync(0xFFFFFFFFu, strip, 0);
                    if (strip >= kStripsPerGroup) break;
                    const int pos0 = strip < kBigStrips ? strip * kStrip : kBigStrips * kStrip + (strip - kBigStrips) * 32;
                    const int run = strip < kBigStrips ? kRun : 1;
                    if (preparing) prep_strip(s, i + 2, pos0, run, lane);  // run batches of 32 positions
                    if (searching) search_strip(s, a.lp, g, i, strip, pos0, run, lane, a.block_shift, probe_all);
                }
            }
            const long long t2 = clock64();
            t_role += t1 - t0;
            t_search += t2 - t1;
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
        __syncthreads();
        const int builder_warp = 31, parser_warp = 24;
        if (a.prof && lane == 0 && (warp == builder_warp || warp == parser_warp || warp == 2 || warp == 3)) {
            // per unit: [0] total, then {role, search} of the builder, the first parser warp, a searcher and a guest
            unsigned long long *o = a.prof + (size_t)unit * 32;
            const int k = warp == builder_warp ? 0 : warp == parser_warp ? 1 : warp == 2 ? 2 : 3;
            if (warp == parser_warp)
                for (int q = 0; q < 6; ++q) o[16 + q] = (unsigned long long)pt[q];
            if (warp == builder_warp) o[0] = (unsigned long long)(clock64() - t_unit0);
            o[1 + 3 * k] = (unsigned long long)t_role;
            o[2 + 3 * k] = (unsigned long long)t_search;
            o[3 + 3 * k] = 0;
        }
    }
}

}  // namespace

LevelParams level_params(int level, int strategy) {
    // chain / nice / good / lazy per level.  The reference only varies the chain length (8/16/32/64 for level
    // <=2/<=4/<=6/else, DS:603-609) and never parses lazily (DS:610); these are tuned with tools/lzmodel.cpp so that
    // level 6 lands just below libz level-6 sizes (4-byte hashes in 16384 buckets make a chain step worth several of the
    // reference's 3-byte ones).
    LevelParams p;
    p.far3 = 256;
    p.far4 = 1024;
    p.min_len = 3;
    p.filtered = strategy == ZK_FILTERED;
    p.skip_deep = 1;
    p.probe_adaptive = level <= 7;
    p.good = 8;
    p.nice = 128;
    p.lazy = 1;
    switch (level) {
        case 1: p.chain = 1; p.lazy = 0; p.nice = 32; break;
        case 2: p.chain = 2; p.lazy = 0; p.nice = 32; break;
        case 3: p.chain = 2; p.nice = 64; break;
        case 4: p.chain = 2; break;
        case 5: p.chain = 3; p.nice = 64; break;
        case 6: p.chain = 3; break;
        case 7: p.chain = 6; break;
        case 8: p.chain = 16; p.good = 32; p.nice = 258; break;
        default: p.chain = 32; p.good = 32; p.nice = 258; p.skip_deep = 0; break;
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
