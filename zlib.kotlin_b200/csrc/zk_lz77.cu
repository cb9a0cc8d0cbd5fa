// zk_lz77.cu -- K1: LZ77 match finding + lazy parse, one CTA per LZ unit.
//
// Replaces the tokeniser of DeflateStream.compressZlibDynamic (deflate/DeflateStream.kt:573-790:
// hash3 DS:588-595, insertTripletAt DS:127-144, chain walk DS:668-692, advanceMatch DS:165-184)
// with a design that maps to an SM instead of a scalar loop.  It is not a port: the reference's
// look-ahead/window arrays, its pre-seeded chains and its stale-window compare do not exist here.
//
// Layout in shared memory (one CTA per SM, ~226 KB):
//   ring  u8 [65536]   input bytes at (q & 0xFFFF); q = unit-relative position, history first
//   prev  u16[32768]   hash-chain links at (q & 32767): previous position with the same 4-byte hash
//   head  u16[16384]   most recent position (q & 0xFFFF) per 14-bit hash
//   match u32[2][4608] per-position best match of the group in flight / the group being parsed (rows of 32 skewed)
//   prep  u8 [2][4096] per position of a group to be built: lower lane of its batch with the same hash
//   tq    u32[32][128] per warp: compare tasks (position, distance) of the strip in flight
//
// Software pipeline over groups of 4096 positions, one __syncthreads per step.  In step i
//   - one elected thread stages group i+4 into the ring with a TMA bulk copy (mbarrier-tracked),
//   - whichever warp takes a strip prepares the same strip of group i+2 (match.any over the hashes of each batch),
//   - warp 31 inserts group i+1 into the hash chains: table updates only, 32 positions per instruction sequence,
//   - warps 16..23 (two per SM sub-partition) parse group i-1: a register-only speculative pass per 16-byte lane
//     segment, a two-level stitch with shuffles, a short real pass from the true entry of every entered segment up to
//     the point where it meets the speculative one, then the tokens leave in order,
//   - every warp (the role warps join when done) steals strips of 128 positions of group i and finds the best match of
//     every position in three phases that keep the lanes busy:
//       walk     a lane follows the hash chains of 4 consecutive positions; a candidate only has to pass a one-byte test
//                (the fourth byte: the hash covers four); candidates that pass become tasks (position, distance) in the
//                warp's queue, their slots handed out with ballots;
//       compare  the warp's lanes take the tasks 32 at a time, whoever produced them: 16 bytes against 16 bytes, the
//                result goes into the match word of the position with a shared-memory atomicMax over (length, nearness),
//                so the order in which tasks are processed does not matter;
//       finish   3-byte matches at small fixed distances where nothing was found (adaptive), exact extension of matches
//                that filled the 16 bytes (cooperative, per distinct distance), the too-far filters.
// Round 1 compared inside the chain loop (three 16-byte compares per position at 16 of 32 lanes active: 17.5 warp
// instructions per input byte); the task queue runs the same compares at full width.
// A search only follows links of candidates within 32768 - 2*4096 bytes, which the builder (one group ahead) can not
// have overwritten, and the best match of a position is a pure function of the data: results do not depend on warp
// timing.  tools/lzmodel2.cpp is a CPU model of exactly this algorithm; tests compare the two token for token.
//
// HBM traffic per input byte: 1 B read (+ 32 KiB halo per unit) and 4 B x tokens/byte written.
// The symbol histograms are built from the token stream by hist_kernel (zk_huffman.cu).
#include "zk_common.cuh"

namespace zk {

namespace {

#ifndef ZK_LZ_THREADS
#define ZK_LZ_THREADS 768
#endif
constexpr int kThreadsLz = ZK_LZ_THREADS;
constexpr int kWarpsLz = kThreadsLz / 32;
constexpr int kHashBits = 14;  // 16384 buckets of 16-bit positions (q & 0xFFFF: the ring holds 64 KiB, links are 16-bit too)
#ifndef ZK_LZ_RUN
#define ZK_LZ_RUN 2
#endif
constexpr int kRun = ZK_LZ_RUN;                     // consecutive positions per lane
constexpr int kStrip = 32 * kRun;            // positions per work-stealing strip
constexpr int kStripsPerGroup = kGroup / kStrip;
constexpr int kTaskCap = kStrip;              // compare tasks a warp queues before it runs them
constexpr int kParserWarps = 8;
constexpr int kFirstParserWarp = kWarpsLz - kParserWarps;  // the last eight warps: two per SM sub-partition
constexpr int kBuilderWarp = 0;
constexpr int kSeg = kGroup / (32 * kParserWarps);  // parser: positions per lane segment
constexpr int kSegs = kGroup / kSeg;
constexpr int kHistGroups = kWindow / kGroup;
constexpr int kMaxDist = kWindow - 2 * kGroup;
constexpr int kStages = 4;                   // ring staging barriers
constexpr int kCap = 16;  // candidates are compared over 16 bytes; the nearest one that fills them is extended
// The short-distance probes (3-byte matches of binary records and runs) cost a fifth of the search and find nothing in
// text: every group probes in its first kProbeSample strips only, and in all of them when the sample of the group
// before it had at least kProbeThreshold hits.  Strips are fixed ranges of positions, so the decision does not depend
// on which warp happens to search them.
constexpr int kProbeSample = 512 / kStrip;  // 512 positions
constexpr int kProbeThreshold = 8;
constexpr uint32_t kFull = 0xFFFFFFFFu;

static_assert(kSeg == 16, "the parser's register-only pass is written for 16-position segments");

struct LzSmem {
    uint32_t ring[65536 / 4 + 8];  // + 32 mirrored bytes so a 20-byte read never wraps
    uint16_t prev[32768];
    uint16_t head[1 << kHashBits];  // most recent position (q & 0xFFFF) per hash
    // best match per position of a group, skewed: position i sits at mvi(i) = i + 4 * (i / 32), so that the parser lanes
    // (one 16-position segment each) read their rows with conflict-free 16-byte loads; + slack: one word past the group.
    // While a strip is being searched: len << 16 | (0xFFFF - dist) (atomicMax); afterwards: len << 16 | (dist - 1).
    uint32_t match[2][kGroup + kGroup / 8 + 8];
    uint8_t prep[2][kGroup];  // per position of a group to be built: 0x80 last of its hash in its batch, 0x40 | lane of
                              // the nearest lower lane with the same hash
    uint32_t tq[kWarpsLz][kTaskCap];  // compare tasks of the strip a warp is searching: (position - group start) << 16 | dist
    uint32_t jmark[kSegs];          // stitch: segments of L's warp the parse visits from L on
    int16_t jnext[kSegs];           // stitch: first segment after L's warp the parse from L leads to
    uint16_t pstart[kSegs];         // entry position (relative to the group) of an entered segment
    unsigned long long bar[kStages];
    int work[2];
    int probe_hits[4];  // short-distance probe hits in the sampled strips of group t, at [t & 3]
    int wcount[kParserWarps];
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

__device__ __forceinline__ int mvi(int i) { return i + ((i >> 5) << 2); }

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
    bool in_aligned;      // the shard's bytes of this unit start on a 16-byte boundary in global memory
};

// Stage group t of the unit into the ring.  Source bytes may come from the shard itself or, for the
// history of unit 0, from the separate history buffer.  Returns true when the TMA path was used.
__device__ __forceinline__ bool stage_group(LzSmem &s, const LzArgs &a, const UnitGeom &g, int t) {
    const int32_t q0 = t * kGroup;
    // the common case -- a whole group of the shard itself, 16-byte aligned in global memory -- is decided with three
    // compares; only the issuing thread does the address arithmetic
    if (q0 >= kWindow && q0 + kGroup <= g.data_end_q && g.in_aligned) {
        if (threadIdx.x == 0) {
            const uint8_t *src = a.in + g.unit_start - kWindow + q0;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&s.bar[t & (kStages - 1)], (uint32_t)kGroup);
            tma_load(reinterpret_cast<uint8_t *>(s.ring) + (q0 & 0xFFFF), src, (uint32_t)kGroup, &s.bar[t & (kStages - 1)]);
        }
        return true;
    }
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
        if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&s.bar[t & (kStages - 1)], (uint32_t)(hi - lo));
            tma_load(ring8 + (q0 & 0xFFFF), src + lo, (uint32_t)(hi - lo), &s.bar[t & (kStages - 1)]);
        }
    } else {
        for (int32_t q = lo + (int32_t)threadIdx.x; q < hi; q += kThreadsLz) ring8[q & 0xFFFF] = src[q];
    }
    return tma;
}

// ---- preparation of a strip of group t, done by whichever warp searches the matching strip ------------------
// For every position: which lower lane of its 32-position batch has the same hash, and whether it is the last of its
// hash in the batch (an exact match.any; cheap here because many warps share its latency, unlike in the serial builder).
// It is a pure function of the batch, so it runs in parallel on the searcher warps and the serial builder is left with
// table updates only.
// match.any answers late (a fifth of all stall samples sat behind it when its result was used at once): a strip's two
// votes are issued before the strip is searched and their results are used after it.
struct PrepVotes {
    uint32_t eq[kStrip / 32];
};
__device__ __forceinline__ PrepVotes prep_issue(const LzSmem &s, int t, int pos0, int lane) {
    PrepVotes v;
    const uint32_t q0 = (uint32_t)t * kGroup + pos0 + lane;
#pragma unroll
    for (int u = 0; u < kStrip / 32; ++u) v.eq[u] = __match_any_sync(kFull, hash4(ring_load32(s.ring, q0 + u * 32)));
    return v;
}
__device__ __forceinline__ void prep_finish(LzSmem &s, int t, int pos0, int lane, const PrepVotes &v) {
    uint8_t *pp = s.prep[t & 1] + pos0;
#pragma unroll
    for (int u = 0; u < kStrip / 32; ++u) {
        const uint32_t eq = v.eq[u];
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
    int32_t data_end_q;
    uint32_t block_shift;
};

// ---- searcher: the best match of every position of a strip ------------------------------------------------------
// Shared-memory accesses of the hot loops by 32-bit shared address: the address arithmetic stays in plain integer
// registers (no generic-to-shared conversion per access) and the loads of independent positions are issued together.
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// 16 bytes at ring offset q (any alignment), ring_a = shared address of the ring
__device__ __forceinline__ Bytes16 ring_load128a(uint32_t ring_a, uint32_t q) {
    const uint32_t wa = ring_a + (q & 0xFFFCu), sh = (q & 3u) << 3;  // words i..i+4; the mirror behind the ring covers the wrap
    const uint32_t a0 = lds_u32(wa), a1 = lds_u32(wa + 4), a2 = lds_u32(wa + 8), a3 = lds_u32(wa + 12), a4 = lds_u32(wa + 16);
    Bytes16 r;
    r.w[0] = __funnelshift_r(a0, a1, sh);
    r.w[1] = __funnelshift_r(a1, a2, sh);
    r.w[2] = __funnelshift_r(a2, a3, sh);
    r.w[3] = __funnelshift_r(a3, a4, sh);
    return r;
}

struct StripAddrs {
    uint32_t ring_a, prev_a, mv_a, tq_a;  // shared addresses: ring, prev, the group's match array, this warp's task queue
};

// The queued compare tasks of a warp: 16 bytes at the position against 16 bytes at the candidate, clipped to the end of
// the deflate block; the longest, on a tie the nearest, stays in the match word (a shared-memory max over
// len << 16 | ~dist: the order in which the tasks are processed does not matter).
__device__ __forceinline__ void run_tasks(const StripAddrs &ad, int T, int32_t g0, int32_t blk_end_q, int lane) {
    __syncwarp();
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int ti = t0 + lane;
        if (ti < T) {
            uint32_t task;  // a volatile load: it stays behind the __syncwarp above
            asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(task) : "r"(ad.tq_a + 4u * (uint32_t)ti) : "memory");
            const uint32_t rel = task >> 16, d = task & 0xFFFFu;
            const uint32_t p = (uint32_t)g0 + rel;
            const int maxlen = min(kMaxMatch, blk_end_q - (int32_t)p);
            const int len = min(common_prefix16(ring_load128a(ad.ring_a, p), ring_load128a(ad.ring_a, p - d)), maxlen);
            if (len >= 4) {
                const uint32_t val = ((uint32_t)len << 16) | (0xFFFFu - d);
                asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(ad.mv_a + 4u * (rel + ((rel >> 5) << 2))), "r"(val) : "memory");
            }
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void search_strip(LzSmem &s, const SearchCtx &cx, int t, int strip, int lane, int warp, bool probe_all) {
    const LevelParams &lp = cx.lp;
    const bool probing = probe_all || strip < kProbeSample;
    uint32_t *mv = s.match[t & 1];
    StripAddrs ad;
    ad.ring_a = smem_addr(s.ring);
    ad.prev_a = smem_addr(s.prev);
    ad.mv_a = smem_addr(mv);
    ad.tq_a = smem_addr(s.tq[warp]);
    const int32_t g0 = t * kGroup;
    const int32_t rel0 = strip * kStrip + lane * kRun;  // the first of this lane's positions, relative to the group
    const int32_t p0 = g0 + rel0;
    // matches never cross the end of their deflate block (a group lies inside one block)
    const int32_t blk_end_q = kWindow + min(cx.ulen, (((g0 - kWindow) >> cx.block_shift) + 1) << cx.block_shift);
    const uint32_t lt = (1u << lane) - 1u;
    int maxlen[kRun];
    uint32_t reach[kRun], c16[kRun], lastd[kRun], need[kRun];
    bool walk[kRun];
    {
        // the links of the four positions (one 8-byte load) and the byte a candidate has to match (index 3)
        if constexpr (kRun == 4) {
            const uint2 lk = *reinterpret_cast<const uint2 *>(&s.prev[p0 & 32767]);
            c16[0] = lk.x & 0xFFFFu;
            c16[1] = lk.x >> 16;
            c16[2] = lk.y & 0xFFFFu;
            c16[3] = lk.y >> 16;
            const uint32_t wa = ad.ring_a + ((uint32_t)p0 & 0xFFFCu);
            const uint32_t w0 = lds_u32(wa), w1 = lds_u32(wa + 4);
            need[0] = w0 >> 24;
            need[1] = w1 & 0xFFu;
            need[2] = (w1 >> 8) & 0xFFu;
            need[3] = (w1 >> 16) & 0xFFu;
        } else {
            const uint32_t lk = *reinterpret_cast<const uint32_t *>(&s.prev[p0 & 32767]);
            c16[0] = lk & 0xFFFFu;
            c16[1] = lk >> 16;
            need[0] = lds_u8(ad.ring_a + ((uint32_t)(p0 + 3) & 0xFFFFu));
            need[1] = lds_u8(ad.ring_a + ((uint32_t)(p0 + 4) & 0xFFFFu));
        }
    }
#pragma unroll
    for (int r = 0; r < kRun; ++r) {
        const int32_t p = p0 + r;
        maxlen[r] = max(0, min(kMaxMatch, blk_end_q - p));
        reach[r] = (uint32_t)min(kMaxDist, max(p - cx.h0q, 0));
        lastd[r] = 0;
        walk[r] = lp.chain > 0 && maxlen[r] >= 4;
    }
    if constexpr (kRun == 4) *reinterpret_cast<uint4 *>(&mv[mvi(rel0)]) = make_uint4(0u, 0u, 0u, 0u);
    else *reinterpret_cast<uint2 *>(&mv[mvi(rel0)]) = make_uint2(0u, 0u);
    // ---- walk: one chain link of each of the four positions per trip, branch-free: the eight loads (next link and test
    // byte of four candidates) are in flight together; candidates that pass the byte test are queued, then compared
    for (int k = 0; k < lp.chain; ++k) {
        bool any_walk = false;
#pragma unroll
        for (int r = 0; r < kRun; ++r) any_walk = any_walk || walk[r];
        if (!__any_sync(kFull, any_walk)) break;
        uint32_t d[kRun], nl[kRun], cb[kRun];
#pragma unroll
        for (int r = 0; r < kRun; ++r) {
            d[r] = ((uint32_t)(p0 + r) - c16[r]) & 0xFFFFu;
            walk[r] = walk[r] && d[r] - 1u < reach[r] && d[r] > lastd[r];
            const uint32_t c = (uint32_t)(p0 + r) - d[r];
            // loaded whether or not the candidate is valid (the addresses are always inside the tables)
            nl[r] = lds_u16(ad.prev_a + ((c & 32767u) << 1));
            cb[r] = lds_u8(ad.ring_a + ((c + 3u) & 0xFFFFu));
        }
        int T = 0;
#pragma unroll
        for (int r = 0; r < kRun; ++r) {
            const bool pass = walk[r] && cb[r] == need[r];
            const uint32_t b = __ballot_sync(kFull, pass);
            if (pass) {
                const uint32_t task = ((uint32_t)(rel0 + r) << 16) | d[r];
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(ad.tq_a + 4u * (uint32_t)(T + __popc(b & lt))), "r"(task) : "memory");
            }
            T += __popc(b);
            if (walk[r]) lastd[r] = d[r];
            c16[r] = nl[r];
        }
        run_tasks(ad, T, g0, blk_end_q, lane);
        if (k + 1 < lp.chain) {
            // a position whose best candidate already fills the 16-byte compare is done (a farther one can only tie, and
            // the nearest wins a tie): long repeats would otherwise queue a task per position and link
#pragma unroll
            for (int r = 0; r < kRun; ++r)
                if ((mv[mvi(rel0) + r] >> 16) >= (uint32_t)kCap) walk[r] = false;
        }
    }
    // ---- finish: probes, extension, filters ----
    int best[kRun], bestd[kRun];
    {
        uint32_t w[kRun];
        if constexpr (kRun == 4) {
            const uint4 mw = *reinterpret_cast<const uint4 *>(&mv[mvi(rel0)]);
            w[0] = mw.x, w[1] = mw.y, w[2] = mw.z, w[3] = mw.w;
        } else {
            const uint2 mw = *reinterpret_cast<const uint2 *>(&mv[mvi(rel0)]);
            w[0] = mw.x, w[1] = mw.y;
        }
#pragma unroll
        for (int r = 0; r < kRun; ++r) {
            best[r] = (int)(w[r] >> 16);
            bestd[r] = best[r] ? (int)(0xFFFFu - (w[r] & 0xFFFFu)) : 0;
        }
    }
    // short matches at small fixed distances (3-byte repeats of binary records, byte runs): the 3 bytes at p against the
    // 3 bytes at p-1, -2, -3, -4, -6, -8, -12, -16, all of which come from the 16 bytes before p
    {
        bool want[kRun];
        bool any_want = false;
#pragma unroll
        for (int r = 0; r < kRun; ++r) {
            want[r] = probing && best[r] < 4 && maxlen[r] >= 3 && lp.min_len <= 3;
            any_want = any_want || want[r];
        }
        if (__any_sync(kFull, any_want)) {
            // the 16 bytes before p0 and the 8 bytes from p0 on: six aligned words
            const uint32_t wi = ((uint32_t)(p0 - 16) >> 2) & 0x3FFFu;
            uint32_t wd[7];
#pragma unroll
            for (int j = 0; j < 6; ++j) wd[j] = s.ring[(wi + j) & 0x3FFFu];
            wd[6] = 0u;
            int hits = 0;
#pragma unroll
            for (int r = 0; r < kRun; ++r) {
                const int sh = 8 * (int)(((uint32_t)p0 & 3u) + r);  // < 32: p0 is a multiple of kRun
                uint32_t bk[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) bk[j] = __funnelshift_r(wd[j], wd[j + 1], sh);
                const uint32_t cw = __funnelshift_r(wd[4], wd[5], sh);
                const uint32_t c0 = cw & 0xFFFFFFu;
                int hit = 0;
                if ((bk[0] & 0xFFFFFFu) == c0) hit = 16;
                if ((bk[1] & 0xFFFFFFu) == c0) hit = 12;
                if ((bk[2] & 0xFFFFFFu) == c0) hit = 8;
                if ((__funnelshift_r(bk[2], bk[3], 16) & 0xFFFFFFu) == c0) hit = 6;
                if ((bk[3] & 0xFFFFFFu) == c0) hit = 4;
                if ((__funnelshift_r(bk[3], cw, 8) & 0xFFFFFFu) == c0) hit = 3;
                if ((__funnelshift_r(bk[3], cw, 16) & 0xFFFFFFu) == c0) hit = 2;
                if ((__funnelshift_r(bk[3], cw, 24) & 0xFFFFFFu) == c0) hit = 1;
                if (want[r] && hit && (uint32_t)hit <= reach[r]) {
                    const uint32_t p = (uint32_t)(p0 + r);
                    const int len = min(common_prefix16(ring_load128(s.ring, p), ring_load128(s.ring, p - (uint32_t)hit)), maxlen[r]);
                    if (len >= 3 && len > best[r]) {
                        best[r] = len;
                        bestd[r] = hit;
                        ++hits;
                    }
                }
            }
            if (strip < kProbeSample) {
                int total = 0;
#pragma unroll
                for (int j = 0; j < kRun; ++j) total += __popc(__ballot_sync(kFull, hits > j));
                if (lane == 0 && total) atomicAdd(&s.probe_hits[t & 3], total);
            }
        }
    }
    // extension of matches that filled the 16-byte compare: cooperative, one round per distinct distance of the strip
    // (both positions of a lane usually share it).  For a distance d the end of every match of the strip is the first x
    // with byte[x] != byte[x-d]; the lanes of the warp compare 3 x 32 words (384 bytes from the strip's first position +
    // 16) and ballot the mismatching words, then each position with that distance finds the first mismatch at or after
    // itself + 16.
    {
        static_assert(kStrip + kCap + kMaxMatch + 4 <= 3 * 32 * 4, "the three ballots cover every match end of the strip");
        bool ext[kRun];
        uint32_t pend[kRun], any_pending = 0u;
#pragma unroll
        for (int r = 0; r < kRun; ++r) {
            ext[r] = best[r] >= kCap && best[r] < maxlen[r];
            pend[r] = __ballot_sync(kFull, ext[r]);
            any_pending |= pend[r];
        }
        const uint32_t scan0 = ((uint32_t)(g0 + strip * kStrip) + kCap) & ~3u;  // word aligned
        while (any_pending) {
            // the distance of the first pending position of the strip
            uint32_t d = 0u;
#pragma unroll
            for (int r = kRun - 1; r >= 0; --r)
                if (pend[r]) d = (uint32_t)__shfl_sync(kFull, bestd[r], __ffs(pend[r]) - 1);
            uint32_t diff[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint32_t x = scan0 + (uint32_t)(k * 32 + lane) * 4u;
                diff[k] = __ballot_sync(kFull, ring_load32(s.ring, x) != ring_load32(s.ring, x - d));
            }
            any_pending = 0u;
#pragma unroll
            for (int r = 0; r < kRun; ++r) {
                const int32_t p = p0 + r;
                const bool mine = ext[r] && (uint32_t)bestd[r] == d;
                if (mine) {
                    // first mismatching word at or after my position + 16
                    const uint32_t from = (uint32_t)p + kCap;  // first byte not yet compared
                    const uint32_t w0 = (from - scan0) >> 2;   // its word index in the bitmap
                    int len = maxlen[r];
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
                        if (xr) {
                            len = (int)(start - (uint32_t)p) + ((__ffs((int)xr) - 1) >> 3);
                        } else {
                            // rare: the mismatch of that word was before `from`; continue byte-exact with a short loop
                            int l2 = (int)(start - (uint32_t)p) + 4;
                            while (l2 < maxlen[r]) {
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
                    best[r] = min(len, maxlen[r]);
                    ext[r] = false;
                }
                pend[r] &= ~__ballot_sync(kFull, mine);
                any_pending |= pend[r];
            }
        }
    }
    uint32_t word[kRun];
#pragma unroll
    for (int r = 0; r < kRun; ++r) {
        int b = best[r];
        if (b == 3 && bestd[r] > lp.far3) b = 0;
        if (b == 4 && bestd[r] > lp.far4) b = 0;
        if (b < lp.min_len || (lp.filtered && b <= 5)) b = 0;
        word[r] = b ? ((uint32_t)b << 16) | (uint32_t)(bestd[r] - 1) : 0u;
    }
    if constexpr (kRun == 4) *reinterpret_cast<uint4 *>(&mv[mvi(rel0)]) = make_uint4(word[0], word[1], word[2], word[3]);
    else *reinterpret_cast<uint2 *>(&mv[mvi(rel0)]) = make_uint2(word[0], word[1]);
}

// ---- parser: eight warps turn the per-position matches of group t into tokens -------------------------------

// The step of the parse at p: 1 for a literal, the match length otherwise (greedy with one-step lazy evaluation).
// (No bounds checks on p + 1 and p + 2: the match words behind the data and the two slack words behind the group are
// zero, and a zero never makes a match give way.)
__device__ __forceinline__ int decide(const uint32_t *mv, int max_lazy, int lazy2, int32_t g0, int32_t p) {
    const int len = (int)(mv[mvi(p - g0)] >> 16);
    const int next = (int)(mv[mvi(p + 1 - g0)] >> 16);
    const int next2 = (int)(mv[mvi(p + 2 - g0)] >> 16);
    const bool defer = ((len < max_lazy) & (next > len)) | ((len < lazy2) & (next2 > len + 1));
    return ((len == 0) | defer) ? 1 : len;
}

__device__ __forceinline__ void parser_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kParserWarps * 32) : "memory"); }

// L = parser lane = segment index.  The carry (first unparsed position, tokens emitted, block start) lives in shared
// memory by group parity: group t reads [t & 1] and writes the other.
template <bool kProf>
__device__ __forceinline__ void parse_group(LzSmem &s, const LzArgs &a, const UnitGeom &g, int t, int L, uint32_t unit_index,
                                            long long (&pt)[6]) {
    long long tc = kProf ? clock64() : 0;
#define ZK_PT(i)                             \
    if (kProf) {                             \
        const long long now = clock64();     \
        pt[i] += now - tc;                   \
        tc = now;                            \
    }
    const LevelParams &lp = a.lp;
    const uint32_t *mv = s.match[t & 1];
    const int lane = L & 31, pw = L >> 5;
    const int32_t g0 = t * kGroup;
    const int32_t gend = min(g0 + kGroup, g.data_end_q);
    const int32_t seg0 = min(g0 + L * kSeg, gend), seg1 = min(seg0 + kSeg, gend);
    // ---- pass A: speculative parse of this lane's segment from its first position, in registers: the 18 match words
    // arrive with independent 16-byte loads; per position the step of the parse (the lazy rule) and, branch-free, "the
    // parse visits position i + step if it visits i".
    const int max_lazy = lp.lazy ? lp.max_lazy : 0, lazy2 = lp.lazy2;
    uint32_t visited = 0u, ismatch = 0u;
    int32_t my_exit = seg0, lastd = -1;
    {
        const int seg_len = seg1 - seg0;
        if (seg_len > 0) {
            const uint32_t *mp = mv + mvi(seg0 - g0);  // the segment's 16 words are contiguous
            const uint4 q0 = *reinterpret_cast<const uint4 *>(mp), q1 = *reinterpret_cast<const uint4 *>(mp + 4),
                        q2 = *reinterpret_cast<const uint4 *>(mp + 8), q3 = *reinterpret_cast<const uint4 *>(mp + 12);
            const uint32_t wa = mv[mvi(seg0 - g0 + kSeg)], wb = mv[mvi(seg0 - g0 + kSeg + 1)];
            const uint32_t w[kSeg + 2] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w, wa, wb};
            int len[kSeg + 2];
#pragma unroll
            for (int i = 0; i < kSeg + 2; ++i) len[i] = (int)(w[i] >> 16);
            uint32_t mm = 0u;
            visited = 1u;
#pragma unroll
            for (int i = 0; i < kSeg; ++i) {
                // bitwise, not short-circuit: the && / || form compiles to a branch per term
                const bool defer = ((len[i] < max_lazy) & (len[i + 1] > len[i])) | ((len[i] < lazy2) & (len[i + 2] > len[i] + 1));
                const bool lit = (len[i] == 0) | defer;
                const int step = lit ? 1 : min(len[i], 16);  // anything from 16 on leaves the segment
                mm |= (lit ? 0u : 1u) << i;
                visited |= ((visited >> i) & 1u) << (i + step);
            }
            visited &= seg_len < kSeg ? (1u << seg_len) - 1u : 0xFFFFu;
            ismatch = visited & mm;
            const int lastpos = 31 - __clz(visited);
            const int step = decide(mv, max_lazy, lazy2, g0, seg0 + lastpos);
            lastd = step > 1 ? (int32_t)(mp[lastpos] & 0xFFFFu) : -1;  // distance - 1
            my_exit = seg0 + lastpos + step;
        }
    }
    // ---- stitch.  The real parse enters the segment that holds the carry, and from an entered segment l the one its
    // speculative exit lies in (the real parse of a segment is made to end at the speculative exit, below): which
    // segments are entered is reachability along those links.  Level 1, inside each parser warp with shuffles (pointer
    // doubling, 5 rounds): for every segment of the warp the set of the warp's segments visited from it (a 32-bit mask)
    // and the first segment after the warp it leads to.  Level 2, after one barrier: every lane walks the hops from
    // warp to warp.
    const int n0 = my_exit >= gend ? kSegs : (my_exit - g0) / kSeg;
    {
        const int wbase = pw * 32;
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
    ZK_PT(0)
    const int par = t & 1;
    const int32_t x0 = s.carry_entry[par];
    const int first = x0 >= gend ? kSegs : (x0 - g0) / kSeg;
    if (L == 0) s.pfinal = x0;  // stays when nothing is entered
    parser_barrier();
    ZK_PT(1)
    bool entered = false;
    {
        int cur = first;
#pragma unroll 1
        while (cur < kSegs) {
            if ((cur >> 5) == pw) entered = (s.jmark[cur] >> lane) & 1u;
            cur = s.jnext[cur];
        }
        if (L == first) s.pstart[L] = (uint16_t)(x0 - g0);
        if (entered) {
            if (n0 < kSegs) s.pstart[n0] = (uint16_t)(my_exit - g0);
            else s.pfinal = my_exit;
        }
    }
    parser_barrier();
    ZK_PT(2)
    // ---- pass B: the real parse of an entered segment from its entry, until it lands on a position the speculative
    // parse visited (usually at once: every literal is such a position).  It may not pass the speculative exit: a
    // match that would is cut (or becomes a literal), and when the two parses never meet the rest up to the exit -- the
    // tail of the speculative parse's last match -- is one match at that distance (or up to two literals).
    uint32_t visited2 = 0u, ismatch2 = 0u;
    int cut_len = 0, gap = 0, m = 0;  // m: first speculative position used, relative to the segment
    if (entered) {
        int32_t p = g0 + (int32_t)s.pstart[L];
        if (p > seg0) {
            bool met = false;
#pragma unroll 1
            while (p < seg1) {
                if ((visited >> (p - seg0)) & 1u) {
                    met = true;
                    break;
                }
                int step = decide(mv, max_lazy, lazy2, g0, p);
                const uint32_t bit = 1u << (p - seg0);
                if (step > 1 && p + step > my_exit) {
                    step = my_exit - p;
                    if (step < 3) step = 1;
                    else cut_len = step;
                }
                visited2 |= bit;
                if (step > 1) ismatch2 |= bit;
                p += step;
            }
            m = met ? p - seg0 : kSeg;
            gap = met ? 0 : my_exit - p;
        }
    }
    const uint32_t spec_mask = m >= 32 ? 0u : visited & ~((1u << m) - 1u);
    int cnt = 0;
    if (entered) cnt = __popc(visited2) + __popc(spec_mask) + (gap >= 3 ? 1 : gap);
    // dense offsets across the parser lanes
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == 31) s.wcount[pw] = incl;
    ZK_PT(3)
    parser_barrier();
    ZK_PT(4)
    int before = 0, total = 0;
#pragma unroll
    for (int k = 0; k < kParserWarps; ++k) {
        const int c = s.wcount[k];
        if (k < pw) before += c;
        total += c;
    }
    const uint32_t tok_off = s.carry_tok_off[par];
    if (entered) {
        uint32_t *dst = a.tokens + (uint64_t)unit_index * a.unit_bytes + tok_off + (uint32_t)(before + incl - cnt);
        // the real prefix covers positions before `m`, the speculative tokens the positions from `m` on: one mask.  The 16
        // match words and the 16 bytes of the segment arrive with five independent vector loads; the stores are
        // predicated straight-line code (a loop over the set bits waits for shared memory once per token)
        const uint32_t vm = visited2 | spec_mask, im = ismatch2 | (ismatch & spec_mask);
        const int last2 = (cut_len && ismatch2) ? 31 - __clz(ismatch2) : -1;
        const uint32_t *mp = mv + mvi(seg0 - g0);
        const uint4 m0 = *reinterpret_cast<const uint4 *>(mp), m1 = *reinterpret_cast<const uint4 *>(mp + 4),
                    m2 = *reinterpret_cast<const uint4 *>(mp + 8), m3 = *reinterpret_cast<const uint4 *>(mp + 12);
        const uint4 rb = *reinterpret_cast<const uint4 *>(s.ring + (((uint32_t)seg0 >> 2) & 0x3FFFu));
        const uint32_t mw[16] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w, m2.x, m2.y, m2.z, m2.w, m3.x, m3.y, m3.z, m3.w};
        const uint32_t bw[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
        for (int i = 0; i < kSeg; ++i) {
            if ((vm >> i) & 1u) {
                uint32_t tok = ((im >> i) & 1u) ? mw[i] : ((bw[i >> 2] >> (8 * (i & 3))) & 0xFFu);
                if (i == last2) tok = ((uint32_t)cut_len << 16) | (tok & 0xFFFFu);
                dst[__popc(vm & ((1u << i) - 1u))] = tok;
            }
        }
        dst += __popc(vm);
        if (gap >= 3) {
            *dst++ = ((uint32_t)gap << 16) | (uint32_t)lastd;
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
    ZK_PT(5)
#undef ZK_PT
}

// kProf: per-role cycle counters for tools/lz_profile.py (a second instantiation: the clock reads are not free)
template <bool kProf>
__global__ void __launch_bounds__(kThreadsLz, 1) lz77_kernel(const LzArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    LzSmem &s = *reinterpret_cast<LzSmem *>(smem_raw);
    // read once through volatile asm: the compiler otherwise re-derives lane and warp from %tid wherever it is short of
    // registers (an S2R plus arithmetic in every inner loop)
    uint32_t tid_once;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_once));
    const int lane = (int)(tid_once & 31u), warp = (int)(tid_once >> 5);
    const bool is_builder = warp == kBuilderWarp;
    const bool is_parser = warp >= kFirstParserWarp && warp < kFirstParserWarp + kParserWarps;
    const int parser_index = warp - kFirstParserWarp;

    for (uint32_t unit = blockIdx.x; unit < a.num_units; unit += gridDim.x) {
        UnitGeom g;
        g.unit_start = (uint64_t)unit * a.unit_bytes;
        g.ulen = (uint32_t)tmin<uint64_t>(a.unit_bytes, a.n - g.unit_start);
        g.h0q = unit == 0 ? (int32_t)(kWindow - a.hist_len) : 0;
        g.data_end_q = kWindow + (int32_t)g.ulen;
        g.first_group = g.h0q / kGroup;
        g.last_group = (g.data_end_q - 1) / kGroup;
        g.in_aligned = ((reinterpret_cast<uintptr_t>(a.in) + g.unit_start) & 15u) == 0;
        SearchCtx cx;
        cx.lp = a.lp;
        cx.h0q = g.h0q;
        cx.ulen = (int32_t)g.ulen;
        cx.data_end_q = g.data_end_q;
        cx.block_shift = a.block_shift;

        // reset tables (deterministic start) and barriers
        {
            uint4 *z = reinterpret_cast<uint4 *>(s.prev);
            const int nvec = (int)((sizeof(s.prev) + sizeof(s.head)) / 16);
            for (int i = threadIdx.x; i < nvec; i += kThreadsLz) z[i] = make_uint4(0, 0, 0, 0);
            // the slack words behind the match arrays are read by the parser's look-ahead at the last position
            if (threadIdx.x < 16) s.match[threadIdx.x >> 3][kGroup + kGroup / 8 + (threadIdx.x & 7)] = 0u;
            if (threadIdx.x == 0) {
                for (int b = 0; b < kStages; ++b) mbar_init(&s.bar[b], 1);
                s.work[0] = s.work[1] = 0;
                s.probe_hits[0] = s.probe_hits[1] = s.probe_hits[2] = s.probe_hits[3] = 0;
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

        // step i: stage(i+4) | prepare(i+2) | build(i+1) | search(i) | parse(i-1)
        long long t_role = 0, t_search = 0;
        long long pt[6] = {0, 0, 0, 0, 0, 0};
        const long long t_unit0 = kProf ? clock64() : 0;
        for (int i = g.first_group - 2; i <= g.last_group + 1; ++i) {
            const bool tma_next = (i + 4 <= g.last_group + 1) ? stage_group(s, a, g, i + 4) : false;
            if (threadIdx.x == 0) {
                s.work[(i + 1) & 1] = 0;
                s.probe_hits[(i + 1) & 3] = 0;
            }
            // (the first group of a unit has no sample before it: it probes everywhere)
            const bool probe_all = !a.lp.probe_adaptive || i <= kHistGroups || s.probe_hits[(i - 1) & 3] >= kProbeThreshold;

            const long long t0 = kProf ? clock64() : 0;
            if (is_builder && i + 1 <= g.last_group && i + 1 >= g.first_group) build_group(s, i + 1, lane);
            if (is_parser && i - 1 >= kHistGroups && i - 1 <= g.last_group)
                parse_group<kProf>(s, a, g, i - 1, parser_index * 32 + lane, unit, pt);
            const long long t1 = kProf ? clock64() : 0;
            const bool searching = i >= kHistGroups && i <= g.last_group;
            const bool preparing = i + 2 >= g.first_group && i + 2 <= g.last_group;
            if (searching || preparing) {
                for (;;) {
                    int strip = 0;
                    // (one lane, one instruction: an add of the constant 1 -- atomicAdd() or atom.add in PTX -- is turned into a
                    // warp-aggregated sequence of twenty instructions by ptxas; inc is not)
                    if (lane == 0) asm volatile("atom.shared.inc.u32 %0, [%1], 0x7FFFFFFF;" : "=r"(strip) : "r"(smem_addr(&s.work[i & 1])) : "memory");
                    strip = __shfl_sync(kFull, strip, 0);
                    if (strip >= kStripsPerGroup) break;
                    PrepVotes votes;
                    if (preparing) votes = prep_issue(s, i + 2, strip * kStrip, lane);
                    if (searching) search_strip(s, cx, i, strip, lane, warp, probe_all);
                    if (preparing) prep_finish(s, i + 2, strip * kStrip, lane, votes);
                }
            }
            const long long t2 = kProf ? clock64() : 0;
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
        if (kProf && a.prof && lane == 0 && (warp == kBuilderWarp || warp == kFirstParserWarp || warp == 2)) {
            // per unit: [0] total; {role, search} cycles of [1,2] the builder, [3,4] the first parser warp, [5,6] a
            // searcher; [8..13] the parser's phases (pass A + stitch 1, wait, stitch 2, pass B + count, wait, copy-out)
            unsigned long long *o = a.prof + (size_t)unit * 32;
            const int k = warp == kBuilderWarp ? 0 : warp == kFirstParserWarp ? 1 : 2;
            if (warp == kBuilderWarp) o[0] = (unsigned long long)(clock64() - t_unit0);
            if (warp == kFirstParserWarp)
                for (int q = 0; q < 6; ++q) o[8 + q] = (unsigned long long)pt[q];
            o[1 + 2 * k] = (unsigned long long)t_role;
            o[2 + 2 * k] = (unsigned long long)t_search;
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
    p.probe_adaptive = level <= 8;
    p.nice = 128;  // (a compare covers 16 bytes; what fills them is extended to its end: nice plays no role any more)
    p.lazy = 1;
    p.max_lazy = 16;
    p.lazy2 = 0;
    // One chain link per position costs about one cycle per byte and SM (profiles/level_sweep_r2.json): level 6 walks
    // one link -- with the second look-ahead it lands within 2 % of libz level 6 on the text corpus and within 0.5 % on
    // the mixed one -- level 7 two (libz level-6 size or better), level 9 eight (smaller than libz level 9).
    switch (level) {
        case 1: p.chain = 1; p.lazy = 0; break;
        case 2: p.chain = 1; p.max_lazy = 4; break;
        case 3: p.chain = 1; p.max_lazy = 8; break;
        case 4: p.chain = 1; break;
        case 5: p.chain = 1; p.lazy2 = 8; break;
        case 6: p.chain = 1; p.lazy2 = 16; p.far4 = 512; break;
        case 7: p.chain = 2; p.lazy2 = 16; p.far4 = 512; break;
        case 8: p.chain = 4; p.lazy2 = 16; p.max_lazy = 32; break;
        default: p.chain = 8; p.lazy2 = 16; p.max_lazy = 258; break;
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
    const uint32_t grid = a.num_units;
    if (prof) {
        ZK_CUDA(cudaFuncSetAttribute(lz77_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LzSmem)));
        lz77_kernel<true><<<grid, kThreadsLz, sizeof(LzSmem), c->stream>>>(a);
    } else {
        ZK_CUDA(cudaFuncSetAttribute(lz77_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LzSmem)));
        lz77_kernel<false><<<grid, kThreadsLz, sizeof(LzSmem), c->stream>>>(a);
    }
    c->launches += 1;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}

}  // namespace zk
