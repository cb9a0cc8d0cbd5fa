// zk_huffman.cu -- K2: per-block code construction and block planning, eight lanes per deflate block (four blocks per warp).
//
// Replaces HuffmanBuilder.buildLengths (deflate/HuffmanBuilder.kt:10-83), CanonicalHuffman.buildEncoder
// (inflate/CanonicalHuffman.kt:85-120), the code-length RLE + header of DeflateStream.kt:818-900 and the
// stored / fixed / dynamic cost choice of DeflateStream.kt:902-955.
//
// Differences from the reference that matter for parity (SURVEY.md section 7 "hard parts"):
//   - lengths come from a real Huffman tree, limited to 15 / 7 bits by moving leaves down on the length
//     counts; the code is always COMPLETE (the reference's Shannon-Fano split can over-subscribe, HB:58-62,
//     and its own inflate rejects incomplete multi-symbol trees, InflateStream.kt:151);
//   - at least two symbols get a code in every tree, so lit[256] > 0 and the distance tree is never empty
//     (InflateStream.kt:102-119);
//   - the literal/length and distance length sequences are run-length coded SEPARATELY, because the
//     reference inflate reads them as two sequences (InflateStream.kt:316-317);
//   - length 258 is symbol 285 (the reference writes 284 + 31, DS:193-203 -- both are valid DEFLATE).
//
// Input:  block_hist[b][320]  (lit/len counts at [0,286), distance counts at [288,318))
// Output: BlockPlan[b]: block type, header bit string, code tables, exact size in bytes.
#include "zk_common.cuh"

namespace zk {

namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kMaxSyms = 288;
// Most of a block's work is sequential (the two-queue tree, the depths, the run-length coding, the header bits) and runs
// on one lane; a warp per block issued those one-lane instruction streams one block at a time, and that -- not latency, not
// occupancy -- bounded the kernel.  So a block gets a GROUP of kG lanes: the parallel parts stride by kG, the sequential ones
// run on the group's first lane, and the four leaders of a warp execute their streams in lockstep.
constexpr int kG = 8;
constexpr int kLitFastBits = 9;            // == zk_inflate.cu kLaneLitBits
constexpr uint32_t kLongSymsPerBlock = 512;  // symbols with longer codes a block may hold before it is coded again, limited
constexpr int kPerWarp = 32 / kG;
__device__ __forceinline__ int group_leader() { return (int)(threadIdx.x & 31u) & ~(kG - 1); }
template <class T>
__device__ __forceinline__ T group_sum(T v) {
#pragma unroll
    for (int off = kG / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
    return v;
}

__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// The kernel is one lane's dependent chain through shared memory for most of its time, so what counts is how many warps
// an SM holds: arrays that are never alive together share their bytes (5 176 instead of 6 904 per warp: 44 instead of 32
// warps per SM).
struct WarpScratch {
    union {
        uint32_t freq[kMaxSyms];          // working copy of the counts (>= 1 for coded symbols): dead once the leaf weights are set
        uint16_t parent[2 * kMaxSyms];    // ... which is when the tree loop starts to write the parents
        uint32_t parent32[16];            // (assign_codes: next code per length)
    };
    uint16_t order[kMaxSyms];     // symbols sorted by ascending (freq, symbol)
    union {
        uint32_t weight[2 * kMaxSyms];    // sort keys, then node weights: dead once the tree stands
        uint8_t depth[2 * kMaxSyms];      // ... which is when the depths are computed
    };
    uint8_t lit_len[kMaxSyms];
    uint8_t dist_len[32];
    uint8_t cl_len[20];
    uint8_t cl_sym[320];          // run-length coded code lengths: symbol
    uint8_t cl_extra[320];        //                              : extra-bit value
    uint32_t cl_freq[20];
    uint32_t cl_code[20];
    int used;
};

__device__ __forceinline__ uint32_t bit_reverse(uint32_t code, int len) { return __brev(code) >> (32 - len); }

// Code lengths from the tree that stands in w.parent (leaves 0 .. used-1 in the order of w.order), limited to maxbits.  One
// lane; no votes or barriers inside, so one group of a warp may call it again with a tighter limit while the others wait.
__device__ void limit_lengths(WarpScratch &w, int used, int maxbits, uint8_t *out) {
    // 4. depths from the root down, counts per length with the overflow folded to maxbits
    const int root = 2 * used - 2;
    w.depth[root] = 0;
    int count[40];
    for (int b = 0; b < 40; ++b) count[b] = 0;
    int overflow = 0;
    for (int i = root - 1; i >= 0; --i) {
        int d = w.depth[w.parent[i]] + 1;
        if (d > 39) d = 39;  // counts beyond maxbits are folded below; keep the byte from wrapping
        w.depth[i] = (uint8_t)d;
        if (i < used) {
            if (d > maxbits) {
                d = maxbits;
                ++overflow;
            }
            count[d]++;
        }
    }
    // 5. repair.  Folding the deeper leaves up to maxbits over-subscribes the code by an exact number of units of
    // 2^-maxbits (the tree was complete).  One step -- a leaf of the deepest level below maxbits that has one becomes an
    // inner node over itself and one of the folded leaves -- takes exactly one unit away: 2^-b turns into 2 * 2^-(b+1),
    // and a leaf leaves level maxbits.  (zlib counts "overflow" over inner nodes as well and gets the same number; counting
    // the leaves alone does not, which only shows when the limit is tight.)
    (void)overflow;
    long long excess = -(1ll << maxbits);
    for (int b = 1; b <= maxbits; ++b) excess += (long long)count[b] << (maxbits - b);
    while (excess > 0 && count[maxbits] > 0) {
        int bits = maxbits - 1;
        while (bits > 0 && count[bits] == 0) --bits;
        if (bits == 0) break;
        count[bits]--;
        count[bits + 1] += 2;
        count[maxbits]--;
        --excess;
    }
    // 6. longest codes to the rarest symbols (order[] is ascending by weight)
    int idx = 0;
    for (int bits = maxbits; bits >= 1; --bits)
        for (int k = count[bits]; k > 0; --k) out[w.order[idx++]] = (uint8_t)bits;
}

// Length-limited Huffman code lengths for freq[0,n).  Group-cooperative where it is cheap (sort by rank),
// the group's first lane for the sequential tree walk.  Every lane returns after the lengths are in `out`.
// lane = lane in the group (0 .. kG-1).
__device__ int build_lengths(WarpScratch &w, const uint32_t *freq_in, int n, int maxbits, uint8_t *out, int lane) {
    // 1. working counts; make sure at least two symbols are coded
    int used = 0;
    for (int i = lane; i < n; i += kG) {
        const uint32_t f = freq_in[i];
        w.freq[i] = f;
        out[i] = 0;
        used += f != 0;
    }
    used = group_sum(used);
    __syncwarp();
    if (lane == 0 && used < 2) {
        for (int i = 0; i < n && used < 2; ++i)
            if (w.freq[i] == 0) {
                w.freq[i] = 1;
                ++used;
            }
    }
    used = __shfl_sync(0xFFFFFFFFu, used, group_leader());
    __syncwarp();
    // 2. rank sort of the coded symbols by (freq, symbol): one key per symbol (freq << 9 | symbol; counts are below
    // 2^22), every lane ranks its symbols, nine at a time, against all keys, which are read as broadcasts in the group
    {
        uint32_t *key = w.weight;  // free until step 3
        for (int i = lane; i < n; i += kG) key[i] = w.freq[i] ? ((w.freq[i] << 9) | (uint32_t)i) : 0xFFFFFFFFu;
        __syncwarp();
        for (int base = 0; base < n; base += 9 * kG) {  // (the trip count depends on n alone: uniform over the warp's groups
            uint32_t mykey[9];                           //  only when they build the same alphabet -- they do)
            int rank[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                mykey[t] = base + lane + kG * t < n ? key[base + lane + kG * t] : 0xFFFFFFFFu;
                rank[t] = 0;
            }
            for (int j = 0; j < n; ++j) {
                const uint32_t kj = key[j];
#pragma unroll
                for (int t = 0; t < 9; ++t) rank[t] += kj < mykey[t];
            }
#pragma unroll
            for (int t = 0; t < 9; ++t)
                if (mykey[t] != 0xFFFFFFFFu) w.order[rank[t]] = (uint16_t)(base + lane + kG * t);
        }
    }
    __syncwarp();
    // 3. tree with two queues (leaves in sorted order, internal nodes in creation order)
    for (int i = lane; i < used; i += kG) w.weight[i] = w.freq[w.order[i]];
    __syncwarp();
    if (lane == 0) {
        int leaf = 0, inner = used, next = used;
        while (next < 2 * used - 1) {
            int pick[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (leaf < used && (inner >= next || w.weight[leaf] <= w.weight[inner])) pick[k] = leaf++;
                else pick[k] = inner++;
            }
            w.weight[next] = w.weight[pick[0]] + w.weight[pick[1]];
            w.parent[pick[0]] = (uint16_t)next;
            w.parent[pick[1]] = (uint16_t)next;
            ++next;
        }
        limit_lengths(w, used, maxbits, out);
    }
    __syncwarp();
    return used;
}

// Canonical codes for `len`, bit-reversed for LSB-first emission; packed as code | len << 16.
__device__ void assign_codes(WarpScratch &w, const uint8_t *len, int n, int maxbits, uint32_t *packed, int lane) {
    // symbols per length with shared atomics, first code of each length by lane 0, then the symbols in chunks of kG in
    // symbol order: a symbol's rank among the codes of its length is the running count of that length plus the number
    // of lower lanes of the chunk with the same length (match.any)
    uint32_t *count = w.cl_freq;   // 16 counters; the code-length frequencies are not needed any more when this runs ...
    uint32_t *next = w.parent32;   // ... and neither is the tree
    for (int i = lane; i < 16; i += kG) count[i] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += kG)
        if (len[i]) atomicAdd(&count[len[i]], 1u);
    __syncwarp();
    if (lane == 0) {
        uint32_t code = 0;
        next[0] = 0;
        for (int b = 1; b <= maxbits; ++b) {
            code = (code + (b > 1 ? count[b - 1] : 0u)) << 1;
            next[b] = code;
        }
    }
    __syncwarp();
    const uint32_t wl = threadIdx.x & 31u;  // lane in the warp: the groups of a warp match among themselves only
    for (int c = 0; c * kG < n; ++c) {
        const int i = c * kG + lane;
        const int l = i < n ? len[i] : 0;
        const uint32_t same = __match_any_sync(0xFFFFFFFFu, (uint32_t)l | ((wl / kG) << 8));
        uint32_t v = 0;
        if (l) v = bit_reverse(next[l] + (uint32_t)__popc(same & ((1u << wl) - 1u)), l) | ((uint32_t)l << 16);
        __syncwarp();
        if (l && (same & ((1u << wl) - 1u)) == 0u) next[l] += (uint32_t)__popc(same);
        if (i < n) packed[i] = v;
        __syncwarp();
    }
}

struct BitSink {
    uint8_t *dst;
    uint32_t nbits;
    uint64_t acc;
    int fill;
    __device__ void put(uint32_t value, int count) {
        acc |= (uint64_t)value << fill;
        fill += count;
        nbits += count;
        while (fill >= 8) {
            *dst++ = (uint8_t)acc;
            acc >>= 8;
            fill -= 8;
        }
    }
    __device__ void flush() {
        if (fill > 0) *dst++ = (uint8_t)acc;
        fill = 0;
        acc = 0;
    }
};

// Run-length code `len[0,n)` into (symbol, extra) pairs appended at w.cl_sym[at..]; returns the new count.
__device__ int rle_lengths(WarpScratch &w, const uint8_t *len, int n, int at) {
    int i = 0;
    while (i < n) {
        const int v = len[i];
        int j = i + 1;
        while (j < n && len[j] == v) ++j;
        int run = j - i;
        i = j;
        if (v == 0) {
            while (run >= 11) {
                const int r = run < 138 ? run : 138;
                w.cl_sym[at] = 18;
                w.cl_extra[at++] = (uint8_t)(r - 11);
                run -= r;
            }
            if (run >= 3) {
                w.cl_sym[at] = 17;
                w.cl_extra[at++] = (uint8_t)(run - 3);
                run = 0;
            }
            while (run-- > 0) {
                w.cl_sym[at] = 0;
                w.cl_extra[at++] = 0;
            }
        } else {
            w.cl_sym[at] = (uint8_t)v;
            w.cl_extra[at++] = 0;
            --run;
            while (run >= 3) {
                const int r = run < 6 ? run : 6;
                w.cl_sym[at] = 16;
                w.cl_extra[at++] = (uint8_t)(r - 3);
                run -= r;
            }
            while (run-- > 0) {
                w.cl_sym[at] = (uint8_t)v;
                w.cl_extra[at++] = 0;
            }
        }
    }
    return at;
}

// ---- per-block symbol histograms from the token stream ------------------------------------------------------
// Replaces the frequency counting of DeflateStream.kt:626-627,726,740-744,755.  One CTA per deflate block; every
// warp counts into its own shared-memory copy (no contention between warps, the hardware aggregates equal addresses
// inside a warp), the copies are summed into block_hist[b][320] with end-of-block added.
constexpr int kHistThreads = 256;
__device__ __forceinline__ int hist_length_symbol(int len) {  // 257..285
    const int y = len - 3;
    if (y < 8) return 257 + y;
    if (len == 258) return 285;
    const int b = 31 - __clz(y);
    return 257 + 4 * (b - 1) + ((y >> (b - 2)) & 3);
}
__device__ __forceinline__ int hist_dist_symbol(int dist_minus1) {  // 0..29
    if (dist_minus1 < 4) return dist_minus1;
    const int b = 31 - __clz(dist_minus1);
    return 2 * b + ((dist_minus1 >> (b - 1)) & 1);
}
__global__ void __launch_bounds__(kHistThreads) hist_kernel(const uint32_t *__restrict__ tokens,
                                                            const uint32_t *__restrict__ block_tok_begin,
                                                            const uint32_t *__restrict__ block_tok_count,
                                                            uint32_t *__restrict__ block_hist) {
    __shared__ uint32_t copy[kHistThreads / 32][kHistStride];
    const int wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (kHistThreads / 32) * kHistStride; i += kHistThreads) (&copy[0][0])[i] = 0;
    __syncthreads();
    const uint32_t blk = blockIdx.x;
    const uint32_t *tok = tokens + block_tok_begin[blk];
    const uint32_t n = block_tok_count[blk];
    uint32_t *mine = copy[wid];
    // four tokens per thread and trip, the loads first: a load followed at once by its use waits for DRAM every time
    for (uint32_t i0 = threadIdx.x; i0 < n; i0 += 4 * kHistThreads) {
        uint32_t t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) t[k] = i0 + k * kHistThreads < n ? __ldg(tok + i0 + k * kHistThreads) : 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (t[k] == 0xFFFFFFFFu) continue;  // not a token: length field 0xFFFF does not occur
            const int len = (int)(t[k] >> 16);
            if (len == 0) {
                atomicAdd(&mine[t[k] & 0xFFu], 1u);
            } else {
                atomicAdd(&mine[hist_length_symbol(len)], 1u);
                atomicAdd(&mine[kDistHistBase + hist_dist_symbol((int)(t[k] & 0xFFFFu))], 1u);
            }
        }
    }
    __syncthreads();
    uint32_t *out = block_hist + (uint64_t)blk * kHistStride;
    for (int i = threadIdx.x; i < kHistStride; i += kHistThreads) {
        uint32_t v = 0;
#pragma unroll
        for (int w = 0; w < kHistThreads / 32; ++w) v += copy[w][i];
        if (i == 256) v += 1;  // end-of-block
        out[i] = v;
    }
}

__device__ __forceinline__ int fixed_litlen_bits(int sym) { return sym < 144 ? 8 : sym < 256 ? 9 : sym < 280 ? 7 : 8; }

__global__ void __launch_bounds__(kWarpsPerCta * 32) huffman_kernel(const uint32_t *__restrict__ block_hist,
                                                                    uint64_t n, uint32_t block_bytes,
                                                                    uint32_t num_blocks, int force_type,
                                                                    BlockPlan *__restrict__ plans) {
    extern __shared__ __align__(16) uint8_t huff_smem[];
    WarpScratch *scratch = reinterpret_cast<WarpScratch *>(huff_smem);
    const int lane = (int)(threadIdx.x & 31u) % kG, wid = threadIdx.x >> 5, sub = (int)(threadIdx.x & 31u) / kG;
    // a group behind the last block repeats the last block (same values to the same places): every lane of the warp has
    // to reach the warp-wide votes and barriers below
    const uint32_t blk = tmin<uint32_t>((blockIdx.x * kWarpsPerCta + wid) * kPerWarp + sub, num_blocks - 1);
    WarpScratch &w = scratch[wid * kPerWarp + sub];
    const uint32_t *hist = block_hist + (uint64_t)blk * kHistStride;
    BlockPlan &plan = plans[blk];

    const int used_lit = build_lengths(w, hist, kNumLitLen, 15, w.lit_len, lane);
    {
        // The inflate side decodes literal/length codes of up to kLaneLitBits = 9 bits from shared memory (zk_inflate.cu); a
        // longer one costs the lane decoder a dependent load from global memory, which the 31 other lanes of its warp wait
        // for.  A text block has a hundred such symbols and nobody notices; a block with a flat literal distribution (binary
        // counters: 3 000 of 18 000 symbols at 10 bits) became the block the whole inflate waited for (6 -> 10 ms per GiB).
        // Such a block gets its lengths from the same tree again, limited to 9 bits, which costs a flat distribution next to
        // nothing (limiting EVERY block costs 1.2 % of the output: text pays for its rare symbols).
        uint32_t long_syms = 0;
        for (int i = lane; i < kNumLitLen; i += kG) long_syms += w.lit_len[i] > kLitFastBits ? hist[i] : 0u;
        long_syms = group_sum(long_syms);
        if (long_syms > kLongSymsPerBlock && lane == 0) limit_lengths(w, used_lit, kLitFastBits, w.lit_len);
        __syncwarp();
    }
    build_lengths(w, hist + kDistHistBase, kNumDist, 15, w.dist_len, lane);

    // header: run-length code the two length sequences separately, then code the RLE alphabet
    int hlit = kNumLitLen, hdist = kNumDist, ncl = 0;
    if (lane == 0) {
        while (hlit > 257 && w.lit_len[hlit - 1] == 0) --hlit;
        while (hdist > 1 && w.dist_len[hdist - 1] == 0) --hdist;
        ncl = rle_lengths(w, w.lit_len, hlit, 0);
        ncl = rle_lengths(w, w.dist_len, hdist, ncl);
        for (int i = 0; i < 19; ++i) w.cl_freq[i] = 0;
        for (int i = 0; i < ncl; ++i) w.cl_freq[w.cl_sym[i]]++;
    }
    hlit = __shfl_sync(0xFFFFFFFFu, hlit, group_leader());
    hdist = __shfl_sync(0xFFFFFFFFu, hdist, group_leader());
    ncl = __shfl_sync(0xFFFFFFFFu, ncl, group_leader());
    __syncwarp();
    build_lengths(w, w.cl_freq, 19, 7, w.cl_len, lane);
    assign_codes(w, w.cl_len, 19, 7, w.cl_code, lane);

    // exact sizes of the three encodings
    uint64_t dyn_bits = 0, fix_bits = 0;
    for (int i = lane; i < kNumLitLen; i += kG) {
        const uint64_t f = hist[i];
        const int extra = i >= 257 ? c_len_extra[i - 257] : 0;
        dyn_bits += f * (uint64_t)(w.lit_len[i] + extra);
        fix_bits += f * (uint64_t)(fixed_litlen_bits(i) + extra);
    }
    for (int i = lane; i < kNumDist; i += kG) {
        const uint64_t f = hist[kDistHistBase + i];
        dyn_bits += f * (uint64_t)(w.dist_len[i] + c_dist_extra[i]);
        fix_bits += f * (uint64_t)(5 + c_dist_extra[i]);
    }
    dyn_bits = group_sum(dyn_bits);
    fix_bits = group_sum(fix_bits);
    int hclen = 19;
    uint32_t head_bits = 0;
    if (lane == 0) {
        while (hclen > 4 && w.cl_len[c_cl_order[hclen - 1]] == 0) --hclen;
        head_bits = 3 + 5 + 5 + 4 + 3 * hclen;
        for (int i = 0; i < ncl; ++i) {
            const int s = w.cl_sym[i];
            head_bits += w.cl_len[s] + (s == 16 ? 2 : s == 17 ? 3 : s == 18 ? 7 : 0);
        }
    }
    hclen = __shfl_sync(0xFFFFFFFFu, hclen, group_leader());
    head_bits = __shfl_sync(0xFFFFFFFFu, head_bits, group_leader());

    const uint32_t raw = (uint32_t)tmin<uint64_t>(block_bytes, n - (uint64_t)blk * block_bytes);
    const uint64_t dyn_total = head_bits + dyn_bits, fix_total = 3 + fix_bits;
    // every block is followed by an empty stored block (3 bits, pad, 00 00 FF FF): the next block starts on a byte
    // and inflate finds one entry point per block (a stored block ends on a byte already; it gets the marker too so
    // that a run of stored blocks is not one long chunk)
    const uint64_t dyn_bytes = ((dyn_total + 3 + 7) >> 3) + 4, fix_bytes = ((fix_total + 3 + 7) >> 3) + 4;
    const uint64_t sto_bytes = 5ull + raw + 5ull;
    int type;
    if (force_type >= 0) type = force_type;
    else if (sto_bytes <= dyn_bytes && sto_bytes <= fix_bytes) type = 0;  // tie order of DS:950-955
    else if (dyn_bytes <= fix_bytes) type = 2;
    else type = 1;
    // a forced encoding must still fit the packer's shared image (block span + header slack)
    if (type != 0 && (type == 2 ? dyn_bytes : fix_bytes) > sto_bytes + 256) type = 0;

    // (every group assigns the dynamic codes -- the votes and barriers inside are the whole warp's -- and a group whose block
    // is not dynamic overwrites them)
    assign_codes(w, w.lit_len, kNumLitLen, 15, plan.litlen_code, lane);
    assign_codes(w, w.dist_len, kNumDist, 15, plan.dist_code, lane);
    if (type != 2) {
        // fixed codes (RFC 1951 3.2.6); also filled for stored blocks so the plan is always defined
        for (int i = lane; i < kNumLitLen; i += kG) {
            const int l = fixed_litlen_bits(i);
            uint32_t code;
            if (i < 144) code = 0x30 + i;
            else if (i < 256) code = 0x190 + (i - 144);
            else if (i < 280) code = i - 256;
            else code = 0xC0 + (i - 280);
            plan.litlen_code[i] = bit_reverse(code, l) | ((uint32_t)l << 16);
        }
        for (int i = lane; i < kNumDist; i += kG) plan.dist_code[i] = bit_reverse((uint32_t)i, 5) | (5u << 16);
    }
    __syncwarp();

    if (lane == 0) {
        BitSink sink{plan.header, 0, 0, 0};
        if (type == 2) {
            sink.put(0, 1);  // BFINAL = 0: the stream is closed by a separate final block
            sink.put(2, 2);
            sink.put((uint32_t)(hlit - 257), 5);
            sink.put((uint32_t)(hdist - 1), 5);
            sink.put((uint32_t)(hclen - 4), 4);
            for (int i = 0; i < hclen; ++i) sink.put(w.cl_len[c_cl_order[i]], 3);
            for (int i = 0; i < ncl; ++i) {
                const int s = w.cl_sym[i];
                sink.put(w.cl_code[s] & 0xFFFFu, (int)(w.cl_code[s] >> 16));
                if (s == 16) sink.put(w.cl_extra[i], 2);
                else if (s == 17) sink.put(w.cl_extra[i], 3);
                else if (s == 18) sink.put(w.cl_extra[i], 7);
            }
        } else if (type == 1) {
            sink.put(0, 1);
            sink.put(1, 2);
        }
        const uint32_t hb = sink.nbits;
        sink.flush();
        plan.type = (uint32_t)type;
        plan.header_bits = hb;
        plan.body_bits = (uint32_t)(type == 2 ? dyn_bits : type == 1 ? fix_bits : 0);
        plan.out_bytes = (uint32_t)(type == 2 ? dyn_bytes : type == 1 ? fix_bytes : sto_bytes);
    }
}

// Exclusive scan of the block sizes -> byte offsets (relative to the body start) and the running total.
// cursor[0] = bytes of body emitted so far (carried across slabs), cursor[1] = error flag.
__global__ void __launch_bounds__(1024) plan_offsets_kernel(const BlockPlan *__restrict__ plans, uint32_t num_blocks,
                                                            uint64_t *__restrict__ offsets, uint64_t *cursor,
                                                            uint64_t cap) {
    __shared__ uint64_t warp_sum[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = cursor[0];
    __syncthreads();
    for (uint32_t base = 0; base < num_blocks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t v = i < num_blocks ? plans[i].out_bytes : 0;
        uint64_t incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
            if ((threadIdx.x & 31) >= off) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint64_t before = 0;
        for (int wdx = 0; wdx < (int)(threadIdx.x >> 5); ++wdx) before += warp_sum[wdx];
        const uint64_t start = carry + before + incl - v;
        if (i < num_blocks) offsets[i] = start;
        __syncthreads();
        if (threadIdx.x == 1023) carry = start + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        cursor[0] = carry;
        if (carry > cap) cursor[1] = 1;  // output does not fit: the packer will not write
    }
}

}  // namespace

int launch_hist(zk_ctx *c, const uint32_t *tokens, const uint32_t *block_tok_begin, const uint32_t *block_tok_count,
                uint32_t num_blocks, uint32_t *block_hist) {
    hist_kernel<<<num_blocks, kHistThreads, 0, c->stream>>>(tokens, block_tok_begin, block_tok_count, block_hist);
    c->launches += 1;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}

int launch_huffman(zk_ctx *c, const uint32_t *block_hist, uint64_t n, uint32_t block_bytes, uint32_t num_blocks,
                   int force_type, BlockPlan *plans) {
    const uint32_t per_cta = kWarpsPerCta * kPerWarp;
    const uint32_t grid = (num_blocks + per_cta - 1) / per_cta;
    const int smem = (int)(sizeof(WarpScratch) * per_cta);
    ZK_CUDA(cudaFuncSetAttribute(huffman_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    huffman_kernel<<<grid, kWarpsPerCta * 32, smem, c->stream>>>(block_hist, n, block_bytes, num_blocks, force_type, plans);
    c->launches += 1;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}

int launch_plan_offsets(zk_ctx *c, const BlockPlan *plans, uint32_t num_blocks, uint64_t *offsets, uint64_t *cursor,
                        uint64_t cap) {
    plan_offsets_kernel<<<1, 1024, 0, c->stream>>>(plans, num_blocks, offsets, cursor, cap);
    c->launches += 1;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}

}  // namespace zk
