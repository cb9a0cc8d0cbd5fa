// zk_inflate.cu -- K5: inflate (zlib / gzip / raw), parallel across blocks and chunks.
//
// Replaces InflateStream.inflateZlib (inflate/InflateStream.kt:376-438) with its block decoders
// (copyStored IS:47-74, decodeFixed IS:265-297, decodeDynamic IS:299-371, readCodeLengths IS:235-263,
// validation IS:102-152), CanonicalHuffman.buildFull/decodeOne (inflate/CanonicalHuffman.kt:54-82,123-138)
// and StreamingBitReader (inflate/StreamingBitReader.kt:9-58).
//
// A DEFLATE stream has no entry points, but streams made by this engine (and by pigz) end every block / chunk with an
// empty stored block: the bytes 00 00 FF FF, after which a block starts on a byte boundary.  The engine
//   1. collects the positions of that pattern (marker_compact_kernel);
//   2. when there are hundreds of them (this engine's own streams): decodes one block per LANE to a token stream
//      (decode_lanes_kernel: tables in shared memory, cp.async input ring, longest blocks first), checks that every
//      block ended exactly on the next marker, prefix-sums the sizes, and expands the tokens with one warp per block to
//      16-bit symbols -- a byte, or 0x8000 | i for "byte i of the 32 KiB that precede this block" (the dictionary of a
//      block is not known yet, so references into it are carried symbolically, copies included); eight blocks are
//      re-based into one coarse chunk (resolve_coarse_kernel);
//      otherwise (pigz chunks, libz / gzip-tool / reference streams): entry points are FOUND -- block_find_kernel accepts
//      the bit positions where a dynamic block header passes every check of the decoder -- and one lane decodes each chunk
//      between two of them (decode_lanes_kernel<true>), a chunk that meets a false entry point going on through its range;
//      only small, damaged or entry-less streams are left to one WARP per marker-delimited chunk (decode_kernel: first to
//      SIZE a chunk, which must end exactly on the next candidate, then to WRITE its symbols), which owns the reference's
//      error codes;
//   3. resolves the last 32 KiB of every chunk: the tail of a chunk is a map over the window before it, maps compose,
//      and the chain becomes a two-level scan (window_compose / window_chain / window_apply kernels);
//   4. turns all symbols into bytes in parallel (materialize_kernel) and checks the trailer (Adler-32 for zlib,
//      CRC-32 + ISIZE for gzip, nothing for raw deflate).
// A preset dictionary is "window -1" in front of the first chunk.  The deflate data may begin inside a byte (begin_bit:
// the stream handle decodes an unfinished stream up to its last entry point and carries the rest).
//
// Error behaviour follows the reference: Z_STREAM_END on success (IS:424), Z_NEED_DICT for FDICT (IS:36-43),
// Z_DATA_ERROR for bad headers / block type 3 / LEN-NLEN / invalid codes / distance too far back
// (IS:287-291) / trailer mismatch, Z_BUF_ERROR for truncated input (IS:425-428: a symbol needs maxLen bits
// available, BR:48-57).  The code-length-code table is built exactly like buildFull, without validation
// (IS:315), so even over-subscribed tables made by the reference's own compressor decode as the reference
// decodes them; the two exceptions that escape inflateZlib (IS:243 -> CH:127,133) become Z_DATA_ERROR.
#include "zk_common.cuh"

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace zk {

namespace {

constexpr int kInfWarps = 4;
constexpr int kLitLutBits = 10;
constexpr int kDistLutBits = 8;
constexpr uint32_t kNoCand = 0xFFFFFFFFu;

__constant__ uint16_t i_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t i_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t i_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t i_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t i_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// Canonical decoder: a direct table for short codes, first-code arithmetic for the long ones.
struct Decoder {
    uint16_t lut[1 << kLitLutBits];  // sym | len << 12 (len 0 = not a short code)
    uint16_t sorted[288];            // symbols in canonical order
    uint16_t first_code[16];         // first canonical code of each length (MSB-first numbering)
    uint16_t first_index[16];        // index into sorted[] of the first symbol of each length
    uint16_t count[16];
    int max_len;
    int lut_bits;
};

struct WarpState {
    Decoder lit, dist;
    uint8_t lengths[320];   // code lengths being read (lit then dist)
    uint8_t cl_bits[128];   // code-length-code full table, built like CanonicalHuffman.buildFull
    uint8_t cl_vals[128];
    int cl_max;
};

struct ChunkInfo {
    uint64_t in_begin;   // byte offset of the chunk's first block
    uint64_t in_end;     // byte offset where it must end (next chunk's start), or stream length for the last
    uint64_t out_len;    // decoded size
    uint64_t out_begin;  // prefix sum of out_len
    uint64_t end_byte;   // byte offset after the chunk's last block (byte-aligned position)
    int32_t status;      // ZK_OK, or the error the reference would report
    int32_t final_seen;  // the chunk ended with a BFINAL block
    int32_t retries;     // times this chunk has absorbed a successor after failing
    int32_t pad;
};

struct BitReader {
    const uint8_t *base;
    uint64_t len;    // stream length in bytes
    uint64_t next;   // next byte to load
    uint64_t hold;
    int held;
    __device__ __forceinline__ void refill() {
        // one aligned 32-bit load per refill once `next` is 4-byte aligned; bytes only at the edges
        if (held <= 32) {
            if ((((uintptr_t)(base + next)) & 3u) == 0 && next + 4 <= len) {
                hold |= (uint64_t)__ldg(reinterpret_cast<const uint32_t *>(base + next)) << held;
                next += 4;
                held += 32;
            } else {
                while (held <= 56 && next < len) {
                    hold |= (uint64_t)__ldg(base + next) << held;
                    ++next;
                    held += 8;
                    if ((((uintptr_t)(base + next)) & 3u) == 0 && held > 24) break;
                }
            }
        }
    }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)(hold & ((1ull << n) - 1ull)); }
    __device__ __forceinline__ void drop(int n) {
        hold >>= n;
        held -= n;
    }
    // position of the next unread bit
    __device__ __forceinline__ uint64_t bit_pos() const { return next * 8 - (uint64_t)held; }
};

// Bit reader of the lane-per-block decoder.  The decode loop is warp-uniform, so a lane that branches away to refill,
// or waits for DRAM, stalls 31 others.  The refill is therefore straight-line predicated code, and the stream reaches
// the lane through a ring of kRing 32-bit words in shared memory filled by cp.async: the word requested at one refill
// is consumed kRing - 1 refills (~30 tokens) later, and no register waits on a load in between (in-order issue would
// stall the warp at the first instruction that touches such a register).  An L2 prefetch runs half a KiB further on.
// A word that is not entirely inside [base, base + len) is read byte by byte (the last word of a raw deflate stream may
// hold its last codes); words behind the stream read as zero.
constexpr int kRing = 8;
struct FastReader {
    uint64_t hold;
    int held;
    uint32_t w0;         // the stream word after the bits in `hold` (ring slot at byte offset roff)
    uint32_t roff;       // byte offset of w0's slot in the lane's ring: slot * 128
    uint32_t ring;       // shared-space address of this lane's slot 0
    const uint8_t *at;   // address of the next word to request (kRing words after w0)
    int32_t words_left;  // whole words of the stream at and after `at`
    uintptr_t lo, hi;    // [base, base + len)
    uint32_t tailw;      // the word that holds the last 1..3 bytes of the stream when it does not end on a word, else 0

    __device__ __forceinline__ uint32_t fetch_edge(uintptr_t a) const {
        if (a >= lo && a + 4 <= hi) return __ldg(reinterpret_cast<const uint32_t *>(a));
        uint32_t w = 0;
        for (int i = 0; i < 4; ++i) {
            const uintptr_t b = a + i;
            if (b >= lo && b < hi) w |= (uint32_t)(*reinterpret_cast<const uint8_t *>(b)) << (i * 8);
        }
        return w;
    }
    __device__ __forceinline__ void seek_bits(uint64_t bitpos) {
        asm volatile("cp.async.wait_all;" ::: "memory");  // nothing may land in the ring after it is rewritten
        const uintptr_t p = lo + (bitpos >> 3);
        const uintptr_t a = p & ~(uintptr_t)3;
        const int sub = (int)(p & 3u) * 8 + (int)(bitpos & 7u);
        hold = (uint64_t)(fetch_edge(a) >> sub);
        held = 32 - sub;
        for (int j = 0; j < kRing; ++j) {
            const uint32_t w = fetch_edge(a + 4 + 4 * j);
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(ring + j * 128), "r"(w) : "memory");
            if (j == 0) w0 = w;
        }
        roff = 0;
        const uintptr_t nxt = a + 4 + 4 * kRing;
        at = reinterpret_cast<const uint8_t *>(nxt);
        const uint64_t left = hi > nxt ? (uint64_t)(hi - nxt) >> 2 : 0;
        words_left = (int32_t)tmin<uint64_t>(left, 0x7FFFFFFFull);
    }
    // move w0 into `hold` when there is room (pred && held <= 32), request the word kRing further on.  Straight-line:
    // every instruction is predicated, the commit / wait pair runs every time (an empty group is free, and the word
    // consumed at a refill was requested at least kRing - 1 commits earlier, which is what wait_group guarantees).
    __device__ __forceinline__ void refill(bool pred) {
        const uint32_t take = (pred && held <= 32) ? 1u : 0u;
        const uint32_t nroff = (roff + 128u) & (uint32_t)((kRing - 1) * 128);
        asm volatile(
            "{\n\t.reg .pred t, p, q, f, z;\n\t.reg .b64 tmp;\n\t.reg .b32 lowat, tw;\n\t"
            "setp.ne.u32 t, %6, 0;\n\t"
            "cvt.u64.u32 tmp, %2;\n\t"
            "shl.b64 tmp, tmp, %1;\n\t"
            "@t or.b64 %0, %0, tmp;\n\t"
            "@t add.s32 %1, %1, 32;\n\t"
            "setp.gt.and.s32 p, %5, 0, t;\n\t"
            "setp.le.and.s32 q, %5, 0, t;\n\t"
            "@p cp.async.ca.shared.global [%7], [%4], 4;\n\t"
            "setp.eq.s32 z, %5, 0;\n\t"
            "selp.b32 tw, %10, 0, z;\n\t"
            "@q st.shared.u32 [%7], tw;\n\t"
            "cp.async.commit_group;\n\t"
            "cp.async.wait_group 7;\n\t"
            "@t ld.shared.u32 %2, [%8];\n\t"
            "@t mov.u32 %3, %9;\n\t"
            "cvt.u32.u64 lowat, %4;\n\t"
            "and.b32 lowat, lowat, 127;\n\t"
            "setp.eq.and.u32 f, lowat, 0, t;\n\t"
            "setp.gt.and.s32 f, %5, 160, f;\n\t"
            "@f prefetch.global.L2 [%4+512];\n\t"
            "@t add.u64 %4, %4, 4;\n\t"
            "@t add.s32 %5, %5, -1;\n\t}"
            : "+l"(hold), "+r"(held), "+r"(w0), "+r"(roff), "+l"(at), "+r"(words_left)
            : "r"(take), "r"(ring + roff), "r"(ring + nroff), "r"(nroff), "r"(tailw)
            : "memory");
    }
    // a raw deflate stream may end inside a word (zlib and gzip streams end with a trailer of four or eight bytes, so their
    // partial word never holds a code): that word is read byte by byte once, and refill() stores it when its turn comes
    __device__ __forceinline__ void set_bounds(const uint8_t *base, uint64_t len) {
        lo = (uintptr_t)base;
        hi = lo + len;
        tailw = (hi & 3u) ? fetch_edge(hi & ~(uintptr_t)3) : 0u;
    }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)hold & ((1u << n) - 1u); }  // n <= 16
    __device__ __forceinline__ void drop(int n) {
        hold >>= n;
        held -= n;
    }
    __device__ __forceinline__ uint64_t bit_pos() const {
        return (uint64_t)((uintptr_t)at - 4 * kRing - lo) * 8 - (uint64_t)held;
    }
};
static_assert(kRing == 8, "cp.async.wait_group immediate in FastReader::refill is kRing - 1");

__device__ __forceinline__ uint32_t rev_bits(uint32_t v, int len) { return __brev(v) >> (32 - len); }

// Build a canonical decoder from code lengths (already validated: complete or single-code).
// kCoop = true: the 32 lanes of a warp work on ONE decoder (lane = lane id).  kCoop = false: the calling thread
// owns the decoder alone (pass lane = 0) -- used when every lane of a warp decodes its own block.
template <bool kCoop>
__device__ __forceinline__ void coop_sync() {
    if (kCoop) __syncwarp();
}
template <bool kCoop>
__device__ void build_decoder(Decoder &d, const uint8_t *lens, int n, int lut_bits, int lane) {
    constexpr int kStep = kCoop ? 32 : 1;
    for (int i = lane; i < 16; i += kStep) d.count[i] = 0;
    for (int i = lane; i < (1 << lut_bits); i += kStep) d.lut[i] = 0;
    coop_sync<kCoop>();
    if (lane == 0) {
        int maxl = 0;
        for (int i = 0; i < n; ++i) {
            const int l = lens[i];
            if (l) {
                d.count[l]++;
                if (l > maxl) maxl = l;
            }
        }
        int code = 0, idx = 0;
        d.first_code[0] = 0;
        d.first_index[0] = 0;
        for (int b = 1; b < 16; ++b) {
            code = (code + (b > 1 ? d.count[b - 1] : 0)) << 1;
            d.first_code[b] = (uint16_t)code;
            d.first_index[b] = (uint16_t)idx;
            idx += d.count[b];
        }
        d.max_len = maxl;
        d.lut_bits = lut_bits;
        // canonical order = by length, then by symbol
        int fill[16];
        for (int b = 0; b < 16; ++b) fill[b] = d.first_index[b];
        for (int i = 0; i < n; ++i) {
            const int l = lens[i];
            if (l) d.sorted[fill[l]++] = (uint16_t)i;
        }
    }
    coop_sync<kCoop>();
    // direct table: every symbol with a short code fills its stride
    const int total = d.first_index[15] + d.count[15];
    for (int k = lane; k < total; k += kStep) {
        const int sym = d.sorted[k];
        const int l = lens[sym];
        if (l > lut_bits) continue;
        const uint32_t code = d.first_code[l] + (uint32_t)(k - d.first_index[l]);
        const uint16_t entry = (uint16_t)(sym | (l << 12));
        for (uint32_t idx = rev_bits(code, l); idx < (1u << lut_bits); idx += 1u << l) d.lut[idx] = entry;
    }
    coop_sync<kCoop>();
}

// Decode one symbol; returns -1 for an invalid prefix.  `br` holds at least max_len bits or the end of
// the stream has been reached (missing bits read as zero; the caller checks availability first).
template <class Reader>
__device__ __forceinline__ int decode_sym(const Decoder &d, Reader &br) {
    const uint32_t e = d.lut[br.peek(d.lut_bits)];
    if (e) {
        br.drop((int)(e >> 12));
        return (int)(e & 0x0FFFu);
    }
    // long code: walk the lengths above the table width
    uint32_t code = rev_bits(br.peek(d.lut_bits), d.lut_bits);
    for (int l = d.lut_bits + 1; l <= d.max_len; ++l) {
        code = (code << 1) | ((uint32_t)(br.hold >> (l - 1)) & 1u);
        const uint32_t rel = code - d.first_code[l];
        if (code >= d.first_code[l] && rel < d.count[l]) {
            br.drop(l);
            return d.sorted[d.first_index[l] + rel];
        }
    }
    return -1;
}

// Read the header of a fixed (type 1) or dynamic (type 2) block and build both decoders.  Returns ZK_OK or the
// code the reference reports (IS:299-348).
template <bool kCoop, class Reader>
__device__ int parse_huffman_header(WarpState &ws, Reader &br, int type, int lane) {
    constexpr int kStep = kCoop ? 32 : 1;
    if (type == 1) {  // IS:154-173
        for (int i = lane; i < 288; i += kStep) ws.lengths[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
        coop_sync<kCoop>();
        build_decoder<kCoop>(ws.lit, ws.lengths, 288, kLitLutBits, lane);
        for (int i = lane; i < 32; i += kStep) ws.lengths[i] = 5;
        coop_sync<kCoop>();
        build_decoder<kCoop>(ws.dist, ws.lengths, 32, kDistLutBits, lane);
    } else {  // IS:299-348
        br.refill();
        if (br.held < 14) {
            return ZK_BUF_ERROR;
        }
        const int hlit = (int)br.peek(5) + 257;
        br.drop(5);
        const int hdist = (int)br.peek(5) + 1;
        br.drop(5);
        const int hclen = (int)br.peek(4) + 4;
        br.drop(4);
        int clen[19];
#pragma unroll
        for (int i = 0; i < 19; ++i) clen[i] = 0;
        int bad = 0;
        for (int i = 0; i < hclen; ++i) {
            br.refill();
            if (br.held < 3) {
                bad = ZK_BUF_ERROR;
                break;
            }
            clen[i_cl_order[i]] = (int)br.peek(3);
            br.drop(3);
        }
        if (bad) {
            return bad;
        }
        // code-length-code table exactly as CanonicalHuffman.buildFull makes it (CH:54-82): later
        // symbols overwrite earlier ones when the lengths over-subscribe the code space
        int cl_max = 0;
        for (int i = 0; i < 19; ++i) cl_max = max(cl_max, clen[i]);
        if (lane == 0) {
            ws.cl_max = cl_max;
            for (int i = 0; i < 128; ++i) ws.cl_bits[i] = 0;
            int per[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nextc[8];
            for (int i = 0; i < 19; ++i)
                if (clen[i]) per[clen[i]]++;
            int code = 0;
            nextc[0] = 0;
            for (int b = 1; b <= 7; ++b) {
                code = (code + per[b - 1]) << 1;
                nextc[b] = code;
            }
            for (int s = 0; s < 19; ++s) {
                const int l = clen[s];
                if (!l) continue;
                const uint32_t assigned = (uint32_t)nextc[l]++;
                for (uint32_t idx = rev_bits(assigned & ((1u << l) - 1u), l); idx < (1u << cl_max); idx += 1u << l) {
                    ws.cl_bits[idx] = (uint8_t)l;
                    ws.cl_vals[idx] = (uint8_t)s;
                }
            }
        }
        coop_sync<kCoop>();
        // IS:235-263, twice: literal/length lengths, then distance lengths, never merged
        int fail = 0;
        for (int pass = 0; pass < 2 && !fail; ++pass) {
            const int count = pass == 0 ? hlit : hdist;
            const int base = pass == 0 ? 0 : 288;
            int i = 0;
            while (i < count) {
                br.refill();
                if (cl_max == 0) {  // "Empty Huffman table": escapes in the reference (CH:127)
                    fail = ZK_DATA_ERROR;
                    break;
                }
                if (br.held < cl_max) {
                    fail = ZK_BUF_ERROR;
                    break;
                }
                const uint32_t look = br.peek(cl_max);
                const int l = ws.cl_bits[look];
                if (l == 0) {  // invalid prefix: escapes in the reference (CH:133)
                    fail = ZK_DATA_ERROR;
                    break;
                }
                const int s = ws.cl_vals[look];
                br.drop(l);
                if (s <= 15) {
                    if (lane == 0) ws.lengths[base + i] = (uint8_t)s;
                    ++i;
                    continue;
                }
                if (s == 16 && i == 0) {  // IS:246: checked before the repeat count is read
                    fail = ZK_DATA_ERROR;
                    break;
                }
                const int nb = s == 16 ? 2 : s == 17 ? 3 : 7;
                br.refill();
                if (br.held < nb) {
                    fail = ZK_BUF_ERROR;
                    break;
                }
                const int rep = (s == 16 ? 3 : s == 17 ? 3 : 11) + (int)br.peek(nb);
                br.drop(nb);
                if (i + rep > count) {
                    fail = ZK_DATA_ERROR;
                    break;
                }
                coop_sync<kCoop>();
                const uint8_t v = s == 16 ? ws.lengths[base + i - 1] : (uint8_t)0;
                coop_sync<kCoop>();
                for (int r = lane; r < rep; r += kStep) ws.lengths[base + i + r] = v;
                i += rep;
            }
            coop_sync<kCoop>();
        }
        if (fail) {
            return fail;
        }
        // IS:102-152: EOB present, a distance code present, neither tree over-subscribed nor
        // incomplete (a single code may be incomplete)
        int ok = 1;
        {
            int any = 0, used = 0, left = 1, maxl = 0;
            int per[16];
            for (int b = 0; b < 16; ++b) per[b] = 0;
            for (int i = 0; i < hlit; ++i) {
                const int l = ws.lengths[i];
                if (l) {
                    any = 1;
                    ++used;
                    per[l]++;
                    maxl = max(maxl, l);
                }
            }
            if (!any || ws.lengths[256] == 0) ok = 0;
            for (int b = 1; b <= maxl && ok; ++b) {
                left = (left << 1) - per[b];
                if (left < 0) ok = 0;
            }
            if (ok && left != 0 && used > 1) ok = 0;
            int dany = 0, dused = 0, dmax = 0;
            for (int b = 0; b < 16; ++b) per[b] = 0;
            for (int i = 0; i < hdist; ++i) {
                const int l = ws.lengths[288 + i];
                if (l) {
                    dany = 1;
                    ++dused;
                    per[l]++;
                    dmax = max(dmax, l);
                }
            }
            if (!dany) ok = 0;
            left = 1;
            for (int b = 1; b <= dmax && ok; ++b) {
                left = (left << 1) - per[b];
                if (left < 0) ok = 0;
            }
            if (ok && left != 0 && dused > 1) ok = 0;
        }
        if (!ok) {
            return ZK_DATA_ERROR;
        }
        build_decoder<kCoop>(ws.lit, ws.lengths, hlit, kLitLutBits, lane);
        build_decoder<kCoop>(ws.dist, ws.lengths + 288, hdist, kDistLutBits, lane);
    }
    return ZK_OK;
}

struct ChunkResult {
    uint64_t out_len;
    uint64_t end_byte;
    int status;
    int final_seen;
};

// Decode the blocks of one chunk.  kWrite = false: size only.  kWrite = true: 16-bit symbols to
// sym[out_begin ...]; references that reach before the chunk become 0x8000 | window index.
// kTokens = true (single-lane mode only): instead of decoded symbols the output is the block's token stream in 16-bit
// units -- a literal is one unit (its byte), a match two (0x8000 | length, distance - 1); `units` counts them.
template <bool kWrite, bool kCoop, bool kTokens = false>
__device__ ChunkResult decode_chunk(WarpState &ws, const uint8_t *in, uint64_t m, uint64_t in_begin, uint64_t in_end,
                                    uint64_t reach_before, uint16_t *sym, uint64_t out_begin, uint64_t out_cap, int lane,
                                    uint32_t *units_out = nullptr, uint32_t first_bit = 0) {
    constexpr int kStep = kCoop ? 32 : 1;
    uint32_t units = 0;
    BitReader br{in, m, in_begin, 0, 0};
    if (first_bit) {  // the data starts inside the first byte (a stream handle that has decoded a prefix)
        br.refill();
        if (br.held >= (int)first_bit) br.drop((int)first_bit);
    }
    uint64_t produced = 0;
    ChunkResult res{0, 0, ZK_OK, 0};
    // every lane runs the same decode; only the copies are split across lanes
    for (;;) {
        br.refill();
        if (br.held < 3) {
            res.status = ZK_BUF_ERROR;
            break;
        }
        const int last = (int)br.peek(1);
        br.drop(1);
        const int type = (int)br.peek(2);
        br.drop(2);
        if (type == 3) {  // IS:406
            res.status = ZK_DATA_ERROR;
            break;
        }
        if (type == 0) {  // IS:47-74
            br.drop(br.held & 7);
            br.refill();
            if (br.held < 32) {
                res.status = ZK_BUF_ERROR;
                break;
            }
            const uint32_t len = br.peek(16);
            br.drop(16);
            const uint32_t nlen = br.peek(16);
            br.drop(16);
            if (((~len) & 0xFFFFu) != nlen) {
                res.status = ZK_DATA_ERROR;
                break;
            }
            // the payload is byte-aligned: step the reader back to a byte position and copy
            const uint64_t src0 = br.next - (uint64_t)(br.held >> 3);
            const uint64_t avail = m - src0;
            const uint32_t take = (uint32_t)tmin<uint64_t>(len, avail);
            if (kWrite) {
                if (produced + take > out_cap) {
                    res.status = ZK_BUF_ERROR;
                    break;
                }
                if (kTokens) {
                    for (uint32_t i = 0; i < take; ++i) sym[out_begin + units + i] = in[src0 + i];
                    units += take;
                } else {
                    for (uint32_t i = lane; i < take; i += kStep) sym[out_begin + produced + i] = in[src0 + i];
                }
                coop_sync<kCoop>();
            }
            produced += take;
            br.next = src0 + take;
            br.hold = 0;
            br.held = 0;
            if (take < len) {  // IS:62-70 runs out of source mid-copy
                res.status = ZK_BUF_ERROR;
                break;
            }
        } else {
            {
                const int hrc = parse_huffman_header<kCoop>(ws, br, type, lane);
                if (hrc != ZK_OK) {
                    res.status = hrc;
                    break;
                }
            }
            // symbol loop, IS:276-295 / IS:350-369
            int err = 0;
            const int lit_need = ws.lit.max_len, dist_need = ws.dist.max_len;
            for (;;) {
                br.refill();
                if (br.held < lit_need) {  // BR:48-57: peek(maxLen) needs that many bits
                    err = ZK_BUF_ERROR;
                    break;
                }
                const int s = decode_sym(ws.lit, br);
                if (s < 0) {
                    err = ZK_DATA_ERROR;
                    break;
                }
                if (s < 256) {
                    if (kWrite) {
                        if (produced >= out_cap) {
                            err = ZK_BUF_ERROR;
                            break;
                        }
                        if (kTokens) sym[out_begin + units++] = (uint16_t)s;
                        else if (lane == 0) sym[out_begin + produced] = (uint16_t)s;
                    }
                    ++produced;
                    continue;
                }
                if (s == 256) break;
                const int lc = s - 257;
                if (lc > 28) {  // IS:195
                    err = ZK_DATA_ERROR;
                    break;
                }
                br.refill();
                const int le = i_len_extra[lc];
                if (br.held < le) {
                    err = ZK_BUF_ERROR;
                    break;
                }
                const int length = i_len_base[lc] + (int)br.peek(le);
                br.drop(le);
                br.refill();
                if (br.held < dist_need) {
                    err = ZK_BUF_ERROR;
                    break;
                }
                const int dsym = decode_sym(ws.dist, br);
                if (dsym < 0 || dsym > 29) {  // IS:207
                    err = ZK_DATA_ERROR;
                    break;
                }
                br.refill();
                const int de = i_dist_extra[dsym];
                if (br.held < de) {
                    err = ZK_BUF_ERROR;
                    break;
                }
                const int dist = i_dist_base[dsym] + (int)br.peek(de);
                br.drop(de);
                // IS:287-291 / 361-365: distance bounded by min(bytes out, 32768).  For a later chunk the
                // bytes before it exist (the predecessor decoded them); the bound is checked when the
                // symbol is resolved.
                // (reach_before: bytes that exist before this chunk -- the preset dictionary for the first one)
                if (dist > kWindow || (uint64_t)dist > produced + reach_before) {
                    err = ZK_DATA_ERROR;
                    break;
                }
                if (kWrite) {
                    if (produced + (uint64_t)length > out_cap) {
                        err = ZK_BUF_ERROR;
                        break;
                    }
                    if (kTokens) {
                        sym[out_begin + units] = (uint16_t)(0x8000u | (uint32_t)length);
                        sym[out_begin + units + 1] = (uint16_t)(dist - 1);
                        units += 2;
                    }
                    coop_sync<kCoop>();
                    uint16_t *dst = sym + out_begin + produced;
                    for (int i = kTokens ? length : lane; i < length; i += kStep) {
                        const int k = dist >= length ? i : i % dist;
                        const int64_t rel = (int64_t)produced - dist + k;  // position inside this chunk
                        uint16_t v;
                        if (rel >= 0) v = sym[out_begin + (uint64_t)rel];
                        else v = (uint16_t)(0x8000u | (uint32_t)(kWindow + rel));
                        dst[i] = v;
                    }
                    coop_sync<kCoop>();
                }
                produced += (uint64_t)length;
            }
            if (err) {
                res.status = err;
                break;
            }
        }
        // block boundary
        const uint64_t bits = br.bit_pos();
        if (last) {
            res.final_seen = 1;
            res.end_byte = (bits + 7) >> 3;
            break;
        }
        if ((bits & 7) == 0) {
            const uint64_t at = bits >> 3;
            if (at == in_end) {
                res.end_byte = at;
                break;
            }
            if (at > in_end) {  // ran past a candidate that was not a real block start
                res.status = ZK_DATA_ERROR;
                break;
            }
        } else if ((bits >> 3) >= in_end) {
            res.status = ZK_DATA_ERROR;
            break;
        }
    }
    res.out_len = produced;
    if (units_out) *units_out = units;
    return res;
}

// ---- kernels --------------------------------------------------------------------------------------------

// zlib header (IS:28-45) -> header[0] = status (ZK_OK / ZK_DATA_ERROR / ZK_NEED_DICT / ZK_BUF_ERROR)
// header[1] = offset of the deflate data (2, or 6 behind a DICTID).  A stream with FDICT needs the caller's dictionary:
// without one the answer is ZK_NEED_DICT (all the reference does, IS:36-43); with one its Adler-32 must be the DICTID.
__global__ void header_kernel(const uint8_t *in, uint64_t m, int32_t *header, int have_dict, uint32_t dict_id) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int st = ZK_OK, body = 2;
    if (m < 2) st = ZK_BUF_ERROR;
    else {
        const int cmf = in[0], flg = in[1];
        if ((cmf & 0x0F) != 8 || ((cmf >> 4) & 0x0F) > 7 || (((cmf << 8) | flg) % 31) != 0) st = ZK_DATA_ERROR;
        else if (flg & 0x20) {
            if (m < 6) st = ZK_BUF_ERROR;
            else if (!have_dict) st = ZK_NEED_DICT;
            else {
                const uint32_t id = ((uint32_t)in[2] << 24) | ((uint32_t)in[3] << 16) | ((uint32_t)in[4] << 8) | (uint32_t)in[5];
                if (id != dict_id) st = ZK_DATA_ERROR;
                body = 6;
            }
        } else if (have_dict) {
            st = ZK_OK;  // zlib.h: a dictionary nobody asked for is ignored
        }
    }
    header[0] = st;
    header[1] = body;
}

// one candidate chunk start per span: the first byte offset p in the span with in[p-4..p) == 00 00 FF FF
__global__ void __launch_bounds__(256) marker_kernel(const uint8_t *__restrict__ in, uint64_t m, uint32_t span_shift,
                                                     uint32_t *__restrict__ cand) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 6; p < m; p += stride) {
        if (in[p - 1] == 0xFF && in[p - 2] == 0xFF && in[p - 3] == 0 && in[p - 4] == 0) {
            // a block must be able to start at p: BTYPE != 3
            if (((in[p] >> 1) & 3) != 3) atomicMin(&cand[p >> span_shift], (uint32_t)p);
        }
    }
}

// candidates -> chunk list (keeps every `every`-th span's candidate); single thread, a few thousand spans
__global__ void chunk_list_kernel(const uint32_t *cand, uint32_t num_spans, uint32_t every, uint64_t m,
                                  ChunkInfo *chunks, uint32_t *num_chunks, uint32_t max_chunks, uint32_t body_begin) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t n = 0;
    chunks[0].in_begin = body_begin;
    n = 1;
    uint32_t since = 0;
    for (uint32_t s = 0; s < num_spans && n < max_chunks; ++s) {
        const uint32_t c = cand[s];
        if (c == kNoCand) continue;
        if (++since < every) continue;
        if ((uint64_t)c <= chunks[n - 1].in_begin) continue;
        chunks[n].in_begin = c;
        ++n;
        since = 0;
    }
    for (uint32_t i = 0; i < n; ++i) {
        chunks[i].in_end = i + 1 < n ? chunks[i + 1].in_begin : m;
        chunks[i].status = ZK_OK;
        chunks[i].final_seen = 0;
        chunks[i].out_len = ~0ull;  // not sized yet
        chunks[i].out_begin = 0;
        chunks[i].end_byte = 0;
        chunks[i].retries = 0;
    }
    num_chunks[0] = n;
}

template <bool kWrite>
__global__ void __launch_bounds__(kInfWarps * 32) decode_kernel(const uint8_t *__restrict__ in, uint64_t m,
                                                                ChunkInfo *chunks, const uint32_t *num_chunks,
                                                                uint16_t *sym, uint64_t out_cap, uint32_t dict_len,
                                                                uint32_t first_bit) {
    __shared__ WarpState state[kInfWarps];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t c = blockIdx.x * kInfWarps + wid;
    if (c >= num_chunks[0]) return;
    ChunkInfo &ci = chunks[c];
    if (!kWrite && ci.out_len != ~0ull) return;  // sized in an earlier round
    const uint64_t room = kWrite ? (ci.out_begin <= out_cap ? out_cap - ci.out_begin : 0) : ~0ull;
    const ChunkResult r = decode_chunk<kWrite, true>(state[wid], in, m, ci.in_begin, ci.in_end, c == 0 ? (uint64_t)dict_len : (uint64_t)kWindow, sym,
                                                     ci.out_begin, room, lane, nullptr, c == 0 ? first_bit : 0u);
    if (lane == 0) {
        ci.status = r.status;
        if (!kWrite) {
            ci.out_len = r.out_len;
            ci.end_byte = r.end_byte;
            ci.final_seen = r.final_seen;
        }
    }
}

// After the size pass: merge chunks whose predecessor did not end on them, prefix-sum the sizes.
// verdict[0] = number of chunks after merging, verdict[1] = 1 if anything was merged (size pass must be
// repeated for the merged chunks), verdict[2] = status of the stream so far, verdict[3..4] = total output.
__global__ void plan_chunks_kernel(ChunkInfo *chunks, uint32_t *num_chunks, uint64_t *verdict) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint32_t n = num_chunks[0];
    uint32_t w = 0;
    bool merged = false, fold_rest = false;
    for (uint32_t i = 0; i < n; ++i) {
        // chunk i is trusted only if the chunk before it ended exactly on its first byte
        if (w > 0 && (fold_rest || chunks[w - 1].status != ZK_OK || chunks[w - 1].final_seen)) {
            if (!fold_rest && chunks[w - 1].final_seen && chunks[w - 1].status == ZK_OK) break;  // stream ended
            // predecessor failed: fold this chunk into it and have it sized again.  A first failure is taken
            // for a false candidate (absorb one successor); a second one for real damage (absorb the rest, so
            // that one warp reports what the reference would report).
            if (!fold_rest && chunks[w - 1].retries >= 1) fold_rest = true;
            chunks[w - 1].in_end = chunks[i].in_end;
            if (!fold_rest) chunks[w - 1].retries += 1;
            chunks[w - 1].status = fold_rest ? ZK_DATA_ERROR : ZK_OK;
            chunks[w - 1].final_seen = 0;
            chunks[w - 1].out_len = ~0ull;  // marks "needs sizing"
            merged = true;
            continue;
        }
        if (w != i) chunks[w] = chunks[i];
        ++w;
    }
    if (fold_rest) chunks[w - 1].status = ZK_OK;
    num_chunks[0] = w;
    uint64_t total = 0;
    int status = ZK_OK;
    for (uint32_t i = 0; i < w; ++i) {
        chunks[i].out_begin = total;
        if (chunks[i].out_len != ~0ull) total += chunks[i].out_len;
        if (chunks[i].status != ZK_OK && i + 1 == w) status = chunks[i].status;
    }
    verdict[0] = w;
    verdict[1] = merged ? 1 : 0;
    verdict[2] = (uint64_t)(int64_t)status;
    verdict[3] = total;
    verdict[4] = w ? chunks[w - 1].end_byte : 0;
    verdict[5] = w ? (uint64_t)chunks[w - 1].final_seen : 0;
}

// Resolve the last 32 KiB of every chunk: windows[c][i] = byte at (chunk c's end - 32768 + i).
// The tail of chunk c is a MAP over the window before it: slot i is a byte, or 0x8000 | (index into the window of
// chunk c-1).  Maps compose (apply the map of c-1 to the references of c), and composition is associative, so the
// chain over thousands of chunks is a two-level scan instead of one serial walk:
//   A. window_compose_kernel: chunks are cut into groups of `group` consecutive chunks; one CTA per group composes
//      its maps in order, so that every map refers to the window BEFORE THE GROUP.  The running map lives in shared
//      memory (ping-pong), the symbols of the next chunk are fetched while the current one is composed.
//   B. window_chain_kernel: one CTA resolves the last window of every group in stream order (n / group steps).
//   C. window_apply_kernel: every other window = its composed map applied to its group's entry window, in parallel.
constexpr int kPerThread = kWindow / 1024;  // window slots per thread, interleaved: thread t owns t, t+1024, ...

// slot (threadIdx.x + k*1024) of the tail map of chunk c, relative to the window of chunk c-1
// Slot (threadIdx.x + k*1024) of the tail map of a chunk, relative to the window of the chunk before it, in two steps so
// that the 32 loads of a thread are issued back to back and nothing waits for them until they are needed: raw symbols
// first (fetch_tail), their interpretation when they are consumed (finish_tail).
__device__ __forceinline__ void fetch_tail(uint64_t begin, uint64_t len, const uint16_t *__restrict__ sym,
                                           uint16_t (&raw)[kPerThread]) {
    const int64_t pos0 = (int64_t)(begin + len) - kWindow + threadIdx.x;
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) {
        const int64_t pos = pos0 + k * 1024;
        raw[k] = sym[pos >= (int64_t)begin ? pos : (int64_t)begin];  // clamped: always a valid address (len >= 0 slack: see host)
    }
}
__device__ __forceinline__ void finish_tail(uint64_t begin, uint64_t len, uint16_t (&sy)[kPerThread], uint32_t *error_flag,
                                            uint32_t dict_len) {
    const int64_t pos0 = (int64_t)(begin + len) - kWindow + threadIdx.x;
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) {
        const int64_t pos = pos0 + k * 1024;
        uint16_t s = 0;
        if (pos >= (int64_t)begin) {
            s = sy[k];
            if ((s & 0x8000u) && (int64_t)begin - kWindow + (int64_t)(s & 0x7FFFu) < -(int64_t)dict_len) {
                atomicExch(error_flag, 1u);  // distance reaches before the start of the output (IS:287-291)
                s = 0;
            }
        } else if (pos >= -(int64_t)dict_len) {
            // the chunk is shorter than the window: the slot shows through to the window before it (for the first
            // chunk that is the preset dictionary, window -1)
            s = (uint16_t)(0x8000u | (uint32_t)((int64_t)kWindow - ((int64_t)begin - pos)));
        }
        sy[k] = s;
    }
}

__global__ void __launch_bounds__(1024) window_compose_kernel(const ChunkInfo *chunks, uint32_t n, uint32_t group,
                                                              const uint16_t *__restrict__ sym, uint16_t *maps,
                                                              uint32_t *error_flag, uint32_t dict_len) {
    extern __shared__ __align__(16) uint16_t msm[];  // two maps of 32768 slots
    const uint32_t c0 = blockIdx.x * group, c1 = min(n, c0 + group);
    uint16_t sy[kPerThread], nx[kPerThread];
    uint64_t begin = chunks[c0].out_begin, len = chunks[c0].out_len;
    fetch_tail(begin, len, sym, sy);
    finish_tail(begin, len, sy, error_flag, dict_len);
    uint64_t nbegin = 0, nlen = 0;
    if (c0 + 1 < c1) nbegin = chunks[c0 + 1].out_begin, nlen = chunks[c0 + 1].out_len;
    for (uint32_t c = c0; c < c1; ++c) {
        // the next chunk's symbols are requested now and interpreted after the barrier; the sizes of the one after it too
        if (c + 1 < c1) fetch_tail(nbegin, nlen, sym, nx);
        uint64_t nnbegin = 0, nnlen = 0;
        if (c + 2 < c1) nnbegin = chunks[c + 2].out_begin, nnlen = chunks[c + 2].out_len;
        const uint16_t *prev = msm + ((c + 1) & 1) * kWindow;
        uint16_t *cur = msm + (c & 1) * kWindow;
        uint16_t *glob = maps + (uint64_t)c * kWindow;
#pragma unroll
        for (int k = 0; k < kPerThread; ++k) {
            const int i = threadIdx.x + k * 1024;
            uint16_t s = sy[k];
            if (c > c0 && (s & 0x8000u)) s = prev[s & 0x7FFFu];
            cur[i] = s;
            glob[i] = s;
        }
        __syncthreads();
        if (c + 1 < c1) {
#pragma unroll
            for (int k = 0; k < kPerThread; ++k) sy[k] = nx[k];
            finish_tail(nbegin, nlen, sy, error_flag, dict_len);
        }
        nbegin = nnbegin, nlen = nnlen;
    }
}

__global__ void __launch_bounds__(1024) window_chain_kernel(uint32_t n, uint32_t group, const uint16_t *__restrict__ maps,
                                                            uint8_t *windows) {
    extern __shared__ __align__(16) uint8_t wsm[];  // two windows of 32 KiB
    const uint32_t groups = (n + group - 1) / group;
    uint16_t sy[kPerThread], nx[kPerThread];
    // window -1 (the preset dictionary right-aligned, zeros otherwise) sits in the slot before windows[0]
    for (int k = 0; k < kPerThread; ++k) wsm[kWindow + threadIdx.x + k * 1024] = (windows - kWindow)[threadIdx.x + k * 1024];
    __syncthreads();
    {
        const uint16_t *mp = maps + (uint64_t)(min(n, group) - 1) * kWindow + threadIdx.x;
#pragma unroll
        for (int k = 0; k < kPerThread; ++k) sy[k] = mp[k * 1024];
    }
    for (uint32_t g = 0; g < groups; ++g) {
        const uint32_t c = min(n, (g + 1) * group) - 1;  // last chunk of the group
        if (g + 1 < groups) {
            const uint16_t *mp = maps + (uint64_t)(min(n, (g + 2) * group) - 1) * kWindow + threadIdx.x;
#pragma unroll
            for (int k = 0; k < kPerThread; ++k) nx[k] = mp[k * 1024];
        }
        const uint8_t *prev = wsm + ((g + 1) & 1) * kWindow;
        uint8_t *cur = wsm + (g & 1) * kWindow;
        uint8_t *glob = windows + (uint64_t)c * kWindow;
#pragma unroll
        for (int k = 0; k < kPerThread; ++k) {
            const int i = threadIdx.x + k * 1024;
            const uint16_t s = sy[k];
            const uint8_t v = (s & 0x8000u) ? prev[s & 0x7FFFu] : (uint8_t)s;
            cur[i] = v;
            glob[i] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kPerThread; ++k) sy[k] = nx[k];
    }
}

__global__ void __launch_bounds__(1024) window_apply_kernel(uint32_t n, uint32_t group, const uint16_t *__restrict__ maps,
                                                            uint8_t *windows) {
    const uint32_t c = blockIdx.x;
    const uint32_t g = c / group;
    if (c >= n || c == min(n, (g + 1) * group) - 1) return;  // group-last windows come from window_chain_kernel
    const uint8_t *entry = windows + ((int64_t)g * group - 1) * kWindow;  // group 0: window -1
    const uint16_t *mp = maps + (uint64_t)c * kWindow;
    uint8_t *glob = windows + (uint64_t)c * kWindow;
#pragma unroll 8
    for (int k = 0; k < kPerThread; ++k) {
        const int i = threadIdx.x + k * 1024;
        const uint16_t s = mp[i];
        glob[i] = (s & 0x8000u) ? entry[s & 0x7FFFu] : (uint8_t)s;
    }
}

// Sharded inflate (one stream cut into chunk ranges across GPUs): the total tail map of a shard relative to the window
// before the shard -- the group-end maps composed in order, in the map domain (a reference stays a reference).
__global__ void __launch_bounds__(1024) window_chain_maps_kernel(uint32_t n, uint32_t group, const uint16_t *__restrict__ maps,
                                                                 uint16_t *__restrict__ total_map) {
    extern __shared__ __align__(16) uint16_t cms[];  // two maps of 32768 slots
    const uint32_t groups = (n + group - 1) / group;
    for (int k = 0; k < kPerThread; ++k) cms[kWindow + threadIdx.x + k * 1024] = (uint16_t)(0x8000u | (threadIdx.x + k * 1024));
    __syncthreads();
    for (uint32_t g = 0; g < groups; ++g) {
        const uint32_t c = min(n, (g + 1) * group) - 1;
        const uint16_t *mp = maps + (uint64_t)c * kWindow;
        const uint16_t *prev = cms + ((g + 1) & 1) * kWindow;
        uint16_t *cur = cms + (g & 1) * kWindow;
        uint16_t sy[kPerThread];
#pragma unroll
        for (int k = 0; k < kPerThread; ++k) sy[k] = mp[threadIdx.x + k * 1024];
#pragma unroll
        for (int k = 0; k < kPerThread; ++k) {
            const uint16_t s = sy[k];
            cur[threadIdx.x + k * 1024] = (s & 0x8000u) ? prev[s & 0x7FFFu] : s;
        }
        __syncthreads();
    }
    const uint16_t *last = cms + ((groups + 1) & 1) * kWindow;  // groups == 0: the identity
    for (int k = 0; k < kPerThread; ++k) total_map[threadIdx.x + k * 1024] = last[threadIdx.x + k * 1024];
}

// The window in front of shard r: the tail maps of shards 0 .. r-1 applied in order to an empty window.
__global__ void __launch_bounds__(1024) window_from_maps_kernel(const uint16_t *__restrict__ maps, uint64_t stride_u16,
                                                                uint32_t count, uint8_t *__restrict__ window) {
    extern __shared__ __align__(16) uint8_t wfm[];  // two windows
    for (int k = 0; k < kPerThread; ++k) wfm[kWindow + threadIdx.x + k * 1024] = 0;
    __syncthreads();
    for (uint32_t q = 0; q < count; ++q) {
        const uint16_t *mp = maps + (uint64_t)q * stride_u16;
        const uint8_t *prev = wfm + ((q + 1) & 1) * kWindow;
        uint8_t *cur = wfm + (q & 1) * kWindow;
        for (int k = 0; k < kPerThread; ++k) {
            const uint16_t s = mp[threadIdx.x + k * 1024];
            cur[threadIdx.x + k * 1024] = (s & 0x8000u) ? prev[s & 0x7FFFu] : (uint8_t)s;
        }
        __syncthreads();
    }
    const uint8_t *last = wfm + ((count + 1) & 1) * kWindow;
    for (int k = 0; k < kPerThread; ++k) window[threadIdx.x + k * 1024] = last[threadIdx.x + k * 1024];
}

// Symbols -> bytes.  One CTA per 16 KiB of output; a thread turns 8 consecutive symbols (one 16-byte load) into one
// 8-byte store.  The chunk of the tile's first position is found once per CTA by binary search, every thread then walks
// forward from it (chunks are almost always much longer than a tile).
__global__ void __launch_bounds__(256) materialize_kernel(const ChunkInfo *__restrict__ chunks,
                                                          const uint32_t *__restrict__ num_chunks,
                                                          const uint16_t *__restrict__ sym,
                                                          const uint8_t *__restrict__ windows, uint8_t *__restrict__ out,
                                                          uint64_t total, uint32_t *error_flag, uint32_t dict_len) {
    __shared__ uint32_t s_lo;
    const uint32_t n = num_chunks[0];
    const uint64_t tile0 = (uint64_t)blockIdx.x * 16384;
    if (threadIdx.x == 0) {
        uint32_t lo = 0, hi = n;  // last chunk with out_begin <= tile0
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (chunks[mid].out_begin <= tile0) lo = mid;
            else hi = mid;
        }
        s_lo = lo;
    }
    __syncthreads();
    uint32_t c = s_lo;
    uint64_t cbegin = chunks[c].out_begin;
    uint64_t nbegin = c + 1 < n ? chunks[c + 1].out_begin : ~0ull;
    const bool aligned8 = (reinterpret_cast<uintptr_t>(out) & 7u) == 0;
    for (int step = 0; step < 8; ++step) {
        const uint64_t p0 = tile0 + (uint64_t)step * 2048 + (uint64_t)threadIdx.x * 8;
        if (p0 >= total) break;
        const int cnt = (int)tmin<uint64_t>(8, total - p0);
        uint32_t w[4] = {0, 0, 0, 0};
        if (cnt == 8) {
            const uint4 v = *reinterpret_cast<const uint4 *>(sym + p0);
            w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
        } else {
            for (int e = 0; e < cnt; ++e) w[e >> 1] |= (uint32_t)sym[p0 + e] << (16 * (e & 1));
        }
        uint64_t bytes = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const uint32_t s16 = (w[e >> 1] >> (16 * (e & 1))) & 0xFFFFu;
            uint32_t v = s16 & 0xFFu;
            if ((s16 & 0x8000u) && e < cnt) {
                const uint64_t p = p0 + e;
                while (nbegin <= p) {  // next chunk (empty chunks are skipped by the loop)
                    ++c;
                    cbegin = nbegin;
                    nbegin = c + 1 < n ? chunks[c + 1].out_begin : ~0ull;
                }
                const uint32_t idx = s16 & 0x7FFFu;
                if ((int64_t)cbegin - kWindow + (int64_t)idx < -(int64_t)dict_len) {
                    atomicExch(error_flag, 1u);
                    v = 0;
                } else {
                    v = windows[((int64_t)c - 1) * kWindow + idx];  // chunk 0: window -1, the preset dictionary
                }
            }
            bytes |= (uint64_t)v << (8 * e);
        }
        if (cnt == 8 && aligned8) {
            *reinterpret_cast<uint64_t *>(out + p0) = bytes;
        } else {
            for (int e = 0; e < cnt; ++e) out[p0 + e] = (uint8_t)(bytes >> (8 * e));
        }
    }
}

// Trailer check and result record.  result = {out_len, consumed, status, checksum}.  container: 0 zlib (big-endian
// Adler-32, IS:414-424), 1 raw deflate (no trailer), 2 gzip (little-endian CRC-32 and ISIZE, RFC 1952).
__global__ void finish_inflate_kernel(const uint8_t *in, uint64_t m, const uint64_t *verdict, const uint32_t *check,
                                      const uint32_t *error_flag, uint64_t *result, int container) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int status = (int)(int64_t)verdict[2];
    const uint64_t total = verdict[3], end_byte = verdict[4];
    uint64_t consumed = end_byte;
    if (status == ZK_OK && error_flag[0]) status = ZK_DATA_ERROR;  // distance too far back (IS:287-291)
    if (status == ZK_OK) {
        const uint64_t need = container == 0 ? 4 : container == 2 ? 8 : 0;
        if (container == 3) {
            // a shard body: it carries no final block and must end exactly on the end of its range
            status = (!verdict[5] && end_byte == m) ? ZK_STREAM_END : ZK_DATA_ERROR;
        } else if (container == 4) {
            // a prefix of a stream without markers: whole chunks up to an entry point; consumed counts BITS
            status = !verdict[5] ? ZK_OK : ZK_DATA_ERROR;
            consumed = verdict[8];
        } else if (!verdict[5]) status = ZK_BUF_ERROR;  // ran out of input before a final block
        else if (end_byte + need > m) status = ZK_BUF_ERROR;  // IS:414-418 trailer bytes missing
        else if (container == 0) {
            const uint32_t t = ((uint32_t)in[end_byte] << 24) | ((uint32_t)in[end_byte + 1] << 16) |
                               ((uint32_t)in[end_byte + 2] << 8) | (uint32_t)in[end_byte + 3];
            consumed = end_byte + 4;
            status = t == check[0] ? ZK_STREAM_END : ZK_DATA_ERROR;  // IS:419-424
        } else if (container == 2) {
            uint32_t t = 0, sz = 0;
            for (int i = 0; i < 4; ++i) {
                t |= (uint32_t)in[end_byte + i] << (8 * i);
                sz |= (uint32_t)in[end_byte + 4 + i] << (8 * i);
            }
            consumed = end_byte + 8;
            status = (t == check[0] && sz == (uint32_t)total) ? ZK_STREAM_END : ZK_DATA_ERROR;
        } else {
            status = ZK_STREAM_END;
        }
    }
    result[0] = total;
    result[1] = consumed;
    result[2] = (uint64_t)(int64_t)status;
    result[3] = check[0];
}

__global__ void fill_u32_kernel(uint32_t *p, uint32_t n, uint32_t v) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}

__global__ void set_result_kernel(uint64_t *result, uint64_t a, uint64_t b, int64_t status, uint64_t d) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        result[0] = a;
        result[1] = b;
        result[2] = (uint64_t)status;
        result[3] = d;
    }
}


// ===== lane-per-block path (streams with many marker-delimited blocks, i.e. this engine's own output) ===========
// The warp-per-chunk decoder above runs the same serial decode on all 32 lanes.  When a stream has thousands of
// blocks of at most 32 KiB, every LANE decodes its own block instead: tables live in global memory (one WarpState
// per block), the latency of a table lookup is hidden by the other lanes and warps, and a warp retires 32 symbols
// per pass through the loop instead of one.  One decode pass: a block writes into its own slot of 32768 symbols, its
// size is known afterwards, and a per-coarse-chunk step moves the slots to their final place while resolving the
// references between the blocks of a coarse chunk; references that reach before the coarse chunk stay symbolic and
// go through the same window chain as above.
constexpr uint32_t kSlotSyms = 32768;   // symbols a block may decode to on this path
constexpr uint32_t kCoarse = 8;         // blocks per coarse chunk
constexpr uint32_t kPieceBytes = 16384; // marker search: bytes per CTA

__device__ __forceinline__ bool marker_at(const uint8_t *in, uint64_t m, uint64_t p) {
    // a block may start at p: the four bytes before it are 00 00 FF FF and BTYPE != 3
    return p >= 6 && p < m && in[p - 1] == 0xFF && in[p - 2] == 0xFF && in[p - 3] == 0 && in[p - 4] == 0 &&
           ((in[p] >> 1) & 3) != 3;
}

// pass 1 (write == false): markers per piece.  pass 2: their positions, in order.
template <bool kWritePass>
__global__ void __launch_bounds__(256) marker_compact_kernel(const uint8_t *__restrict__ in, uint64_t m,
                                                             uint32_t *__restrict__ counts,
                                                             const uint64_t *__restrict__ piece_offset,
                                                             uint64_t *__restrict__ positions, uint64_t max_positions) {
    __shared__ uint32_t warp_tot[8];
    const uint64_t base = (uint64_t)blockIdx.x * kPieceBytes + (uint64_t)threadIdx.x * (kPieceBytes / 256);
    uint64_t found = 0;  // bit k: marker at base + k (64 positions per thread)
    if (base >= 16 && base + 80 <= m && (reinterpret_cast<uintptr_t>(in) & 15u) == 0) {
        // bytes [base - 16, base + 80) as 24 words; position base + k needs the four bytes before it and its own
        const uint4 *v = reinterpret_cast<const uint4 *>(in + base - 16);
        uint32_t w[24];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const uint4 x = __ldg(v + q);
            w[4 * q] = x.x, w[4 * q + 1] = x.y, w[4 * q + 2] = x.z, w[4 * q + 3] = x.w;
        }
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            const int o = 12 + k;  // window offset of the byte at base + k - 4
            const uint32_t four = __funnelshift_r(w[o >> 2], w[(o >> 2) + 1], (o & 3) * 8);
            const uint32_t self = (w[(o + 4) >> 2] >> (((o + 4) & 3) * 8)) & 0xFFu;
            if (four == 0xFFFF0000u && ((self >> 1) & 3u) != 3u) found |= 1ull << k;
        }
    } else {
        for (int k = 0; k < (int)(kPieceBytes / 256); ++k)
            if (marker_at(in, m, base + k)) found |= 1ull << k;
    }
    const uint32_t mine = (uint32_t)__popcll(found);
    uint32_t incl = mine;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (int w = 0; w < 8; ++w) {
        if (w < wid) before += warp_tot[w];
        total += warp_tot[w];
    }
    if (!kWritePass) {
        if (threadIdx.x == 0) counts[blockIdx.x] = total;
    } else {
        uint64_t at = piece_offset[blockIdx.x] + before + incl - mine;
        while (found) {
            const int k = __ffsll((long long)found) - 1;
            found &= found - 1;
            if (at < max_positions) positions[at] = base + (uint64_t)k;
            ++at;
        }
    }
}

// exclusive scan of the piece counts (single CTA); total to out_total[0]
__global__ void __launch_bounds__(1024) marker_scan_kernel(const uint32_t *__restrict__ counts, uint32_t n,
                                                           uint64_t *__restrict__ offsets, uint64_t *out_total) {
    __shared__ uint64_t warp_sum[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t v = i < n ? counts[i] : 0;
        uint64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint64_t before = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += warp_sum[w];
        const uint64_t start = carry + before + incl - v;
        if (i < n) offsets[i] = start;
        __syncthreads();
        if (threadIdx.x == 1023) carry = start + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) out_total[0] = carry;
}

struct LaneChunk {
    uint64_t in_begin, in_end;  // bytes (the bytes that hold begin_bits / target_bits)
    uint64_t out_begin;
    uint32_t out_len;
    int32_t status;
    uint64_t end_byte;
    int32_t final_seen;
    uint32_t units;  // 16-bit units of the block's token stream
    uint64_t stored_src;  // a chunk that is one stored block: byte offset of its payload in the stream ...
    uint32_t stored_len;  // ... and its length (then units == 0)
    uint32_t dead;        // set by plan_lanes_kernel: the chunk started at a false entry point, the chunk before it covers it
    uint64_t begin_bits;   // bit position of the chunk's first block (a multiple of 8 behind a sync marker)
    uint64_t target_bits;  // where it must end: the next chunk's begin_bits, or the end of the stream
    uint64_t end_bits;     // where its last block did end
    uint64_t slot_off;     // first unit of its token stream in the slot array
    // a chunk of a foreign stream that ran past its target (the entry point behind it was false) goes on into the next
    // chunk's range; the units of that part are in the spill pool
    uint64_t slot_off_b;   // first unit of the second part in the spill pool
    uint32_t units_a;      // units of the first part (== units when the chunk stayed inside its range)
    uint32_t ext;          // 1: the chunk covers the next chunk's range as well, and that chunk is dead
};
// a chunk of a foreign stream whose decode passed target_bits inside a block: the entry point behind it was not one
constexpr int32_t kLaneOverrun = 1000;

__global__ void lane_table_kernel(const uint64_t *__restrict__ positions, uint32_t num_markers, uint64_t m,
                                  LaneChunk *__restrict__ lc, uint32_t body_begin, uint32_t first_bit) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > num_markers) return;
    lc[k].in_begin = k == 0 ? body_begin : positions[k - 1];
    lc[k].in_end = k < num_markers ? positions[k] : m;
    lc[k].begin_bits = lc[k].in_begin * 8 + (k == 0 ? first_bit : 0u);
    lc[k].target_bits = lc[k].in_end * 8;
    lc[k].end_bits = 0;
    lc[k].slot_off = (uint64_t)k * kSlotSyms;
    lc[k].dead = 0;
    lc[k].slot_off_b = 0;
    lc[k].units_a = 0;
    lc[k].ext = 0;
    lc[k].out_begin = 0;
    lc[k].out_len = 0;
    lc[k].status = ZK_OK;
    lc[k].end_byte = 0;
    lc[k].final_seen = 0;
    lc[k].units = 0;
    lc[k].stored_src = 0;
    lc[k].stored_len = 0;
}

// one LANE per block.  The loop is warp-uniform: in every pass each lane that is still busy either reads one block
// header (and builds its tables) or decodes exactly one token, then the lanes meet again at the vote -- without that
// the lanes of a warp would run their data-dependent loops one after the other.
// Token sink of the lane decoder.  A 2-byte store per token from 32 lanes is 32 sectors through the L1 tag stage, and
// every shared-memory load of the warp queues behind it; tokens are therefore gathered four at a time in a register pair
// and leave as one 8-byte store.
struct TokenSink {
    uint64_t acc;
    uint32_t cnt;    // units in acc (0..3)
    uint16_t *ptr;   // next 8-byte aligned group
    __device__ __forceinline__ void put(uint32_t v, uint32_t n) {  // n = 1: v = unit; n = 2: v = unit0 | unit1 << 16
        acc |= (uint64_t)v << (16u * cnt);
        const uint32_t spill = (cnt + n > 4u) ? (v >> 16) : 0u;  // cnt == 3 and n == 2: the second unit fell off the top
        cnt += n;
        if (cnt >= 4u) {
            *reinterpret_cast<uint64_t *>(ptr) = acc;
            ptr += 4;
            acc = spill;
            cnt -= 4u;
        }
    }
    __device__ __forceinline__ uint32_t finish(const uint16_t *base) {  // -> units written in total
        for (uint32_t i = 0; i < cnt; ++i) ptr[i] = (uint16_t)(acc >> (16u * i));
        return (uint32_t)(ptr - base) + cnt;
    }
};

// Codes longer than the shared-memory table of the lane decoder: six consecutive lengths (L0 .. L0+5) decode by canonical
// arithmetic from six registers -- per length the left-justified 15-bit upper limit of its codes and (first index in the
// canonical order - first code) -- plus ONE load from the decoder's symbol order.  The walk through the decoder in global
// memory (several dependent loads) is left for lengths beyond that.
struct LongCodes {
    uint32_t lim[3];  // lim(L0 + 2i) | lim(L0 + 2i + 1) << 16, in units of 2^-15 (the last limit of a full code is 32768)
    uint32_t off[3];  // the same packing of (first_index - first_code) as int16
    __device__ void load(const Decoder &d, int l0) {
        for (int i = 0; i < 3; ++i) {
            uint32_t lm = 0, of = 0;
            for (int h = 0; h < 2; ++h) {
                const int l = l0 + 2 * i + h;
                uint32_t v = 0, o = 0;
                if (l <= 15) {
                    v = ((uint32_t)d.first_code[l] + (uint32_t)d.count[l]) << (15 - l);
                    o = ((uint32_t)d.first_index[l] - (uint32_t)d.first_code[l]) & 0xFFFFu;
                }
                if (v > 0xFFFFu) v = 0xFFFFu;  // 65535 > every 15-bit code: same verdicts as 32768
                lm |= v << (16 * h);
                of |= o << (16 * h);
            }
            lim[i] = lm;
            off[i] = of;
        }
    }
    // bits = the next (at least 15) stream bits, LSB first.  -> index into d.sorted and the code length, or len = 0
    __device__ __forceinline__ void find(uint32_t bits, int l0, int &index, int &len) const {
        const uint32_t c15 = __brev(bits) >> 17;
        const uint32_t l_0 = lim[0] & 0xFFFFu, l_1 = lim[0] >> 16, l_2 = lim[1] & 0xFFFFu, l_3 = lim[1] >> 16,
                       l_4 = lim[2] & 0xFFFFu, l_5 = lim[2] >> 16;
        const int k = (c15 >= l_0) + (c15 >= l_1) + (c15 >= l_2) + (c15 >= l_3) + (c15 >= l_4) + (c15 >= l_5);
        const uint32_t pair = k < 2 ? off[0] : k < 4 ? off[1] : off[2];
        const int o = (int)(int16_t)((k & 1) ? (pair >> 16) : (pair & 0xFFFFu));
        const int l = l0 + k;
        len = (k < 6 && l <= 15) ? l : 0;
        index = (int)(c15 >> (15 - min(l, 15))) + o;
    }
};

// ===== entry points into streams somebody else made ==========================================================
// A stream written by libz (or by the reference) has no sync markers, but most of its blocks are dynamic-Huffman blocks,
// and the header of such a block is very unlike random bits: BFINAL/BTYPE, HLIT and HDIST in range, a COMPLETE code-length
// code (Kraft sum exactly 1), code lengths that decode without overflow, a literal/length code that is complete and has an
// end-of-block code, a distance code that is complete or a single code.  Those are the checks parse_huffman_header applies
// (IS:299-348, 102-152) -- an accepted position is a header the decoder will take -- plus the complete code-length code.
// Measured on libz output (tools/blockfind_model.cpp, the CPU model of this kernel): 12.5 % of all bit positions pass the
// type test, 0.045 % have a complete code-length code as well, and of 16 M positions per file every accepted one was a true
// block start.  block_find_kernel: one CTA per span of 8 KiB; the FIRST accepted bit position of the span is a candidate
// chunk start.  Per tile of 1 KiB every thread tests its 32 bit positions for the type / HLIT / HDIST pattern and queues the
// hits (at most every third position can be one); the queue is worked off with the Kraft sum (seven lookups in a 512-entry
// table of three 3-bit lengths); what passes is queued again, and once a warp's worth has gathered warp 0 decodes those
// headers completely, one per lane.  A false candidate costs a second decode of the chunk before it (run_inflate merges
// the two), a missed block start only its entry point.
constexpr uint32_t kFindSpan = 8192;
constexpr uint64_t kForeignMinBytes = 256 << 10;  // below that the warp-per-chunk decoder is as quick
constexpr uint64_t kForeignMinChunks = 8;
constexpr int kForeignRounds = 8;                 // merges of consecutive false entry points before the slow path takes over
constexpr uint32_t kFindSlack = 1024;   // bytes staged behind the span: a header is at most ~560 bytes
constexpr uint32_t kFindTile = 1024;    // bytes per tile: 32 bit positions per thread
constexpr int kFindThreads = 256;
constexpr uint32_t kFindSub = 4;        // candidates per span: the first accepted position of every quarter (2 KiB)
constexpr uint32_t kFindSubBits = kFindSpan * 8 / kFindSub;
constexpr uint32_t kFindQ1 = kFindTile * 8 / 3 + 8, kFindQ2 = kFindTile * 8 / 3 + 64;
constexpr uint64_t kNoBits = ~0ull;

struct FindSmem {
    uint32_t data[(kFindSpan + kFindSlack) / 4 + 8];
    uint32_t q2[kFindQ2];   // bit positions in the span whose code-length code is complete
    uint16_t q1[kFindQ1];   // bit positions in the tile that look like BFINAL 0, BTYPE 2, HLIT <= 29, HDIST <= 29
    uint8_t kraft3[512];
    uint8_t pre[32][128];   // warp 0: every lane's table of the code-length code (length << 5 | symbol)
    uint32_t warp_tot[kFindThreads / 32];
    uint32_t n1, n2;
    uint32_t best[kFindSub];
};

// 32 bits of the stream at bit position `bit` of the span: from the staged bytes, or from global memory behind them; bits
// past the end of the stream read as zero
__device__ __forceinline__ uint32_t find_bits32(const FindSmem &sm, const uint8_t *__restrict__ in, uint64_t m, uint64_t span_byte,
                                                uint32_t bit) {
    const uint32_t w = bit >> 5, sh = bit & 31;
    if ((w + 2) * 4 <= kFindSpan + kFindSlack) return __funnelshift_r(sm.data[w], sm.data[w + 1], sh);
    uint64_t v = 0;
    const uint64_t b0 = span_byte + (uint64_t)w * 4;
    for (int i = 0; i < 8; ++i)
        if (b0 + i < m) v |= (uint64_t)in[b0 + i] << (8 * i);
    return (uint32_t)(v >> sh);
}

// the rest of the checks for one position whose code-length code is complete.  `tab` = this lane's 128-entry table.
// Everything but that table lives in registers: the 19 lengths by static index, the next canonical code of each length
// as a byte of a 64-bit word, the two trees as their Kraft sums (2^15 = complete; the per-length "over-subscribed" test of
// IS:102-152 is the same as sum <= 2^15, because the partial sums only grow).
__device__ bool find_check_lengths(const FindSmem &sm, const uint8_t *__restrict__ in, uint64_t m, uint64_t span_byte, uint32_t at,
                                   uint8_t *tab) {
    const uint64_t nbits = m * 8, abs0 = span_byte * 8;
    const uint32_t h = find_bits32(sm, in, m, span_byte, at);
    const int hlit = (int)((h >> 3) & 31u) + 257, hdist = (int)((h >> 8) & 31u) + 1, hclen = (int)((h >> 13) & 15u) + 4;
    uint32_t pos = at + 17;
    {
        constexpr int ord[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        const uint32_t f0 = find_bits32(sm, in, m, span_byte, pos), f1 = find_bits32(sm, in, m, span_byte, pos + 30);
        int clen[19];
        uint64_t count = 0;  // byte l: symbols of length l
#pragma unroll
        for (int i = 0; i < 19; ++i) {
            const int l = i < hclen ? (int)((i < 10 ? f0 >> (3 * i) : f1 >> (3 * (i - 10))) & 7u) : 0;
            clen[ord[i]] = l;
            count += 1ull << (8 * l);
        }
        pos += 3 * hclen;
        uint64_t next = 0;  // byte l: next canonical code of length l
        uint32_t code = 0;
#pragma unroll
        for (int b = 1; b <= 7; ++b) {
            code = (code + (b > 1 ? (uint32_t)(count >> (8 * (b - 1))) & 0xFFu : 0u)) << 1;
            next |= (uint64_t)(code & 0xFFu) << (8 * b);
        }
#pragma unroll
        for (int sym = 0; sym < 19; ++sym) {
            const int l = clen[sym];
            if (!l) continue;
            const uint32_t c = (uint32_t)(next >> (8 * l)) & 0xFFu;
            next += 1ull << (8 * l);
            for (uint32_t idx = rev_bits(c, l); idx < 128u; idx += 1u << l) tab[idx] = (uint8_t)(l << 5 | sym);
        }
    }
    uint32_t kraft0 = 0, kraft1 = 0;
    int used0 = 0, used1 = 0, eob = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const int count = pass == 0 ? hlit : hdist;
        uint32_t kraft = 0;
        int i = 0, prev = 0, used = 0;
        while (i < count) {
            if (abs0 + pos + 1 > nbits) return false;
            const uint32_t look = find_bits32(sm, in, m, span_byte, pos);
            const uint32_t e = tab[look & 127u];
            const int sym = (int)(e & 31u);
            pos += e >> 5;
            if (sym <= 15) {
                if (sym) {
                    kraft += 32768u >> sym;
                    ++used;
                    if (pass == 0 && i == 256) eob = 1;
                }
                prev = sym;
                ++i;
                continue;
            }
            if (sym == 16 && i == 0) return false;  // IS:246
            const int nb = sym == 16 ? 2 : sym == 17 ? 3 : 7;
            const int rep = (sym == 18 ? 11 : 3) + (int)((look >> (e >> 5)) & ((1u << nb) - 1u));
            pos += nb;
            if (i + rep > count) return false;
            const int v = sym == 16 ? prev : 0;
            if (v) {
                kraft += (uint32_t)rep * (32768u >> v);
                used += rep;
                if (pass == 0 && i <= 256 && 256 < i + rep) eob = 1;
            }
            prev = v;
            i += rep;
        }
        if (pass == 0) {
            kraft0 = kraft;
            used0 = used;
        } else {
            kraft1 = kraft;
            used1 = used;
        }
    }
    if (abs0 + pos > nbits) return false;
    if (!eob || !used1) return false;  // IS:102-152: an end-of-block code and a distance code
    if (kraft0 > 32768u || (kraft0 != 32768u && used0 > 1)) return false;
    if (kraft1 > 32768u || (kraft1 != 32768u && used1 > 1)) return false;
    return true;
}

// cand[span * 4 + q] = the first accepted bit position of quarter q of the span behind min_bits (kNoBits if there is none),
// flags[...] = 1 / 0.  Quarters, so that streams of small blocks (a flush every few KiB, pigz pieces of very compressible
// data) do not end up with eight blocks, eight header parses, behind one entry point.
__global__ void __launch_bounds__(kFindThreads, 4) block_find_kernel(const uint8_t *__restrict__ in, uint64_t m, uint64_t min_bits,
                                                                uint64_t *__restrict__ cand, uint32_t *__restrict__ flags) {
    extern __shared__ __align__(16) uint8_t find_smem_raw[];
    FindSmem &sm = *reinterpret_cast<FindSmem *>(find_smem_raw);
    const uint64_t span_byte = (uint64_t)blockIdx.x * kFindSpan;
    const uint64_t nbits = m * 8, span_bits0 = span_byte * 8;
    const int tid = threadIdx.x;
    {
        // stage the span and its slack; bytes past the stream are zero
        const uint32_t words = (kFindSpan + kFindSlack) / 4 + 8;
        const bool aligned = (reinterpret_cast<uintptr_t>(in) & 3u) == 0;
        for (uint32_t w = tid; w < words; w += kFindThreads) {
            const uint64_t b = span_byte + (uint64_t)w * 4;
            uint32_t v = 0;
            if (aligned && b + 4 <= m) {
                v = __ldg(reinterpret_cast<const uint32_t *>(in + b));
            } else {
                for (int i = 0; i < 4; ++i)
                    if (b + i < m) v |= (uint32_t)in[b + i] << (8 * i);
            }
            sm.data[w] = v;
        }
        for (int i = tid; i < 512; i += kFindThreads) {  // Kraft weight of three 3-bit lengths: 128 >> l for l != 0
            int k = 0;
            for (int f = 0; f < 3; ++f) {
                const int l = (i >> (3 * f)) & 7;
                if (l) k += 128 >> l;
            }
            sm.kraft3[i] = (uint8_t)k;
        }
        if (tid == 0) {
            sm.n1 = 0;
            sm.n2 = 0;
            for (uint32_t q = 0; q < kFindSub; ++q) sm.best[q] = 0xFFFFFFFFu;
        }
    }
    __syncthreads();
    const uint32_t span_bytes = (uint32_t)tmin<uint64_t>(kFindSpan, m > span_byte ? m - span_byte : 0);
    for (uint32_t tile = 0; tile < span_bytes; tile += kFindTile) {
        {
            // type / HLIT / HDIST at 32 positions per thread (its own word and the next): a mask of hits, a scan of their
            // counts over the CTA, the positions written in order
            const uint32_t w = tile / 4 + tid;
            const uint32_t w0 = sm.data[w], w1 = sm.data[w + 1];
            // all 32 positions at once, a bit per position: bit j of (W >> k) is stream bit j + k.  BFINAL 0 and BTYPE 10b are
            // bits (0, 0, 1); a 5-bit field is 30 or 31 when its upper four bits are set
            const uint64_t W = ((uint64_t)w1 << 32) | w0;
            const uint64_t type = ~W & ~(W >> 1) & (W >> 2);
            const uint64_t hlit_bad = (W >> 4) & (W >> 5) & (W >> 6) & (W >> 7);
            const uint64_t hdist_bad = (W >> 9) & (W >> 10) & (W >> 11) & (W >> 12);
            uint32_t hits = (uint32_t)(type & ~hlit_bad & ~hdist_bad);
            const uint32_t mine = (uint32_t)__popc(hits);
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if ((tid & 31) >= o) incl += v;
            }
            if ((tid & 31) == 31) sm.warp_tot[tid >> 5] = incl;
            __syncthreads();
            uint32_t before = 0, total = 0;
#pragma unroll
            for (int q = 0; q < kFindThreads / 32; ++q) {
                const uint32_t t = sm.warp_tot[q];
                if (q < (tid >> 5)) before += t;
                total += t;
            }
            uint32_t at = before + incl - mine;
            while (hits) {
                const int j = __ffs(hits) - 1;
                hits &= hits - 1;
                sm.q1[at++] = (uint16_t)(tid * 32 + j);
            }
            if (tid == 0) sm.n1 = total;
        }
        __syncthreads();
        const uint32_t n1 = sm.n1;
        for (uint32_t i = tid; i < n1; i += kFindThreads) {
            // complete code-length code: HCLEN + 4 lengths of 3 bits from bit 17 on, 57 bits at most
            const uint32_t at = tile * 8 + sm.q1[i];
            const uint32_t w = at >> 5, sh = at & 31;
            const uint32_t a = sm.data[w], b = sm.data[w + 1], c2 = sm.data[w + 2], d = sm.data[w + 3];
            const uint32_t x0 = __funnelshift_r(a, b, sh), x1 = __funnelshift_r(b, c2, sh), x2 = __funnelshift_r(c2, d, sh);
            const int hclen = (int)((x0 >> 13) & 15u) + 4;
            uint64_t f = ((uint64_t)__funnelshift_r(x1, x2, 17) << 32) | __funnelshift_r(x0, x1, 17);
            f &= (1ull << (3 * hclen)) - 1ull;
            int k = 0;
#pragma unroll
            for (int t = 0; t < 7; ++t) k += sm.kraft3[(uint32_t)(f >> (9 * t)) & 511u];
            if (k == 128 && span_bits0 + at + 33 <= nbits && span_bits0 + at > min_bits) sm.q2[atomicAdd(&sm.n2, 1u)] = at;
        }
        __syncthreads();
        const uint32_t n2 = sm.n2;
        const bool last_tile = tile + kFindTile >= span_bytes;
        if (n2 >= 32u || (last_tile && n2)) {
            if (tid < 32) {
                for (uint32_t i = tid; i < n2; i += 32) {
                    const uint32_t at = sm.q2[i];
                    if (find_check_lengths(sm, in, m, span_byte, at, sm.pre[tid])) atomicMin(&sm.best[at / kFindSubBits], at);
                }
            }
            __syncthreads();
            if (tid == 0) sm.n2 = 0;
        }
        __syncthreads();
    }
    if (tid < (int)kFindSub) {
        const uint32_t best = sm.best[tid];
        cand[(uint64_t)blockIdx.x * kFindSub + tid] = best != 0xFFFFFFFFu ? span_bits0 + best : kNoBits;
        flags[(uint64_t)blockIdx.x * kFindSub + tid] = best != 0xFFFFFFFFu ? 1u : 0u;
    }
}

// candidates -> chunk table: chunk 0 starts at the first block of the stream, chunk k at the k-th candidate
__global__ void foreign_table_kernel(const uint64_t *__restrict__ cand, const uint32_t *__restrict__ flags,
                                     const uint64_t *__restrict__ off, uint32_t spans, uint64_t total, uint64_t m,
                                     uint64_t first_bits, LaneChunk *__restrict__ lc) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > spans) return;
    uint64_t k, begin;
    if (s == spans) {  // chunk 0
        k = 0;
        begin = first_bits;
    } else {
        if (!flags[s]) return;
        k = off[s] + 1;
        begin = cand[s];
        lc[k - 1].target_bits = begin;
        lc[k - 1].in_end = begin >> 3;
    }
    lc[k].in_begin = begin >> 3;
    lc[k].begin_bits = begin;
    if (k == total) {
        lc[k].target_bits = m * 8;
        lc[k].in_end = m;
    }
    lc[k].out_begin = 0;
    lc[k].out_len = 0;
    lc[k].status = ZK_OK;
    lc[k].end_byte = 0;
    lc[k].end_bits = 0;
    lc[k].final_seen = 0;
    lc[k].units = 0;
    lc[k].stored_src = 0;
    lc[k].stored_len = 0;
    lc[k].slot_off = (begin + 3) & ~3ull;  // a unit per bit position, four-unit groups aligned
    lc[k].dead = 0;
    lc[k].slot_off_b = 0;
    lc[k].units_a = 0;
    lc[k].ext = 0;
}

// Longest blocks first: the decode of a block is serial, so the kernel can not end before its longest block does, and a
// warp waits for the slowest of its 32 lanes.  Blocks are bucketed by compressed size (a proxy for their token count)
// in descending order; a warp then takes 32 blocks of similar size, the big ones start at once.
__global__ void __launch_bounds__(1024) lane_order_kernel(const LaneChunk *__restrict__ lc, uint32_t n, uint32_t *perm) {
    __shared__ uint32_t bucket[256];
    if (threadIdx.x < 256) bucket[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < n; k += 1024) {
        const uint64_t size = lc[k].in_end - lc[k].in_begin;
        atomicAdd(&bucket[255u - (uint32_t)tmin<uint64_t>(size >> 7, 255)], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int b = 0; b < 256; ++b) {
            const uint32_t c = bucket[b];
            bucket[b] = run;
            run += c;
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < n; k += 1024) {
        const uint64_t size = lc[k].in_end - lc[k].in_begin;
        perm[atomicAdd(&bucket[255u - (uint32_t)tmin<uint64_t>(size >> 7, 255)], 1u)] = k;
    }
}

constexpr int kLaneLitBits = 9;   // literal/length codes up to 9 bits and ...
constexpr int kLaneDistBits = 6;  // ... distance codes up to 6 bits decode from shared memory
constexpr int kLaneWarps = 6;     // 6 x 37 KiB of tables and rings: one persistent CTA per SM, 192 blocks in flight on it
constexpr int kLaneLitEntries = 1 << kLaneLitBits, kLaneDistEntries = 1 << kLaneDistBits;
constexpr int kLaneWarpSmem = (kLaneLitEntries + kLaneDistEntries) * 32 * 2 + kRing * 32 * 4;
// kForeign = false: the blocks of this engine's own streams (one block per lane, at most kSlotSyms symbols, a stored block
// only alone).  kForeign = true: chunks of a stream somebody else made, cut at the dynamic block headers block_find_kernel
// found: any number of blocks of any type and size per lane, positions in bits.  Such a chunk's tokens go to the slot array
// at the chunk's own bit position: a literal costs at least one bit and one unit, a match at least two bits and two units,
// a stored block at least 35 bits and five units, and a chunk starts with a dynamic header of 40 bits or more -- so a
// chunk's units stay below the bit position its decode has reached, and chunks never meet.  A token is only written when
// the position behind it is not past the chunk's target: a lane that runs into the next chunk (the entry point was false)
// stops with kLaneOverrun before it touches that chunk's units, and is decoded again with the chunks merged.
template <bool kForeign>
__global__ void __launch_bounds__(kLaneWarps * 32, 1) decode_lanes_kernel(const uint8_t *__restrict__ in, uint64_t m,
                                                                          LaneChunk *lc, uint32_t num_chunks,
                                                                          WarpState *tables, uint16_t *slots,
                                                                          uint32_t *next_task,
                                                                          const uint32_t *__restrict__ perm, uint32_t dict_len,
                                                                          uint16_t *spill, unsigned long long *spill_next,
                                                                          uint64_t spill_units) {
    // every lane's short-code tables, interleaved so that the lanes of a warp fall into different banks: entry i of
    // lane l sits at [i * 32 + l].  Longer codes go through the lane's WarpState in global memory.
    extern __shared__ __align__(16) uint16_t s_lut_all[];
    uint16_t *s_lit = s_lut_all + (threadIdx.x >> 5) * (kLaneWarpSmem / 2);
    uint16_t *s_dist = s_lit + kLaneLitEntries * 32;
    const int lane = threadIdx.x & 31;
    const uint32_t num_tasks = (num_chunks + 31) / 32;
    for (;;) {
        uint32_t task = 0;
        if (lane == 0) task = atomicAdd(next_task, 1u);
        task = __shfl_sync(0xFFFFFFFFu, task, 0);
        if (task >= num_tasks) break;
        const bool live = task * 32 + lane < num_chunks;
        const uint32_t k = live ? perm[task * 32 + lane] : 0u;
        enum { kHeader = 0, kSymbols = 1, kDone = 2 };
        int state = live ? kHeader : kDone;
        int status = ZK_OK, final_seen = 0, last = 0;
        uint64_t end_byte = 0, end_bits = 0;
        uint32_t produced = 0, stored_len = 0;
        uint64_t stored_src = 0;
        const uint32_t kk = live ? k : 0;
        WarpState &ws = tables[kk];
        uint16_t *out = slots + lc[kk].slot_off;
        TokenSink sink{0ull, 0u, out};
        uint64_t target = lc[kk].target_bits;
        // foreign streams: the chunk has passed its target inside a block (the entry point there was false).  It goes on
        // through that chunk's range -- once -- with its further units in the spill pool, so nothing of the other lane's is
        // touched; -> false when that is not possible (second time, last chunk, pool full): kLaneOverrun, decoded again.
        bool extended = false;
        uint32_t units_a = 0;
        uint64_t slot_b = 0;
        auto extend = [&]() -> bool {
            if (!kForeign || extended || target >= m * 8) return false;
            const uint64_t target2 = lc[kk + 1].target_bits;
            const unsigned long long need = ((target2 - target) + 64ull + 3ull) & ~3ull;  // a unit per bit, as in the slot array
            const unsigned long long at = atomicAdd(spill_next, need);
            if (at + need > spill_units) return false;
            units_a = sink.finish(out);
            slot_b = at;
            out = spill + at;
            sink = TokenSink{0ull, 0u, out};
            target = target2;
            extended = true;
            return true;
        };
        constexpr uint32_t kOutCap = kForeign ? 0x7FFF0000u : kSlotSyms;  // symbols a chunk may decode to
        // bytes that exist before the block: the preset dictionary for block 0 (IS:287-291), a full window otherwise
        const uint32_t reach_before = (live && k == 0) ? dict_len : (uint32_t)kWindow;
        LongCodes long_lit, long_dist;
        FastReader fr;
        fr.ring = (uint32_t)__cvta_generic_to_shared(s_dist + kLaneDistEntries * 32) + lane * 4;
        fr.set_bounds(in, m);
        fr.seek_bits(lc[kk].begin_bits);
        while (__any_sync(0xFFFFFFFFu, state != kDone)) {
            if (state == kHeader) {
                // once or twice per block: the plain byte-wise reader parses the header, the fast one is re-seated after
                const uint64_t bp = fr.bit_pos();
                BitReader br{in, m, bp >> 3, 0, 0};
                br.refill();
                if (br.held < 3 + (int)(bp & 7)) {
                    status = ZK_BUF_ERROR;
                    state = kDone;
                } else {
                    br.drop((int)(bp & 7));
                    last = (int)br.peek(1);
                    br.drop(1);
                    const int type = (int)br.peek(2);
                    br.drop(2);
                    if (type == 3) {
                        status = ZK_DATA_ERROR;
                        state = kDone;
                    } else if (type == 0) {
                        br.drop(br.held & 7);
                        br.refill();
                        if (br.held < 32) {
                            status = ZK_BUF_ERROR;
                            state = kDone;
                        } else {
                            const uint32_t len = br.peek(16);
                            br.drop(16);
                            const uint32_t nlen = br.peek(16);
                            br.drop(16);
                            const uint64_t src0 = br.next - (uint64_t)(br.held >> 3);
                            if (((~len) & 0xFFFFu) != nlen) {
                                status = ZK_DATA_ERROR;
                                state = kDone;
                            } else if (src0 + len > m || produced + len > kOutCap ||
                                       (!kForeign && len && (produced || stored_len))) {
                                // truncated, more than this path's slot holds, or (this engine's streams) a chunk that
                                // mixes a stored payload with other data: the warp-per-chunk path takes the stream
                                status = ZK_BUF_ERROR;
                                state = kDone;
                            } else if (kForeign && (src0 + len) * 8 > target && !(extend() && (src0 + len) * 8 <= target)) {
                                status = kLaneOverrun;
                                state = kDone;
                            } else {
                                if (len) {  // the payload is copied by the expansion kernel, in parallel
                                    if (kForeign) {
                                        // five units: 0x8000 (a match head of length 0), then 45 + 3 bits in units that are
                                        // skipped like the distance of a match -- the source offset and the length
                                        const uint32_t s32 = (uint32_t)src0;
                                        sink.put(0x8000u | ((0x8000u | (s32 & 0x7FFFu)) << 16), 2u);
                                        sink.put((0x8000u | ((s32 >> 15) & 0x7FFFu)) | ((0x8000u | (s32 >> 30) | ((len & 0x1FFFu) << 2)) << 16), 2u);
                                        sink.put(len >> 13, 1u);
                                    } else {
                                        stored_src = src0;
                                        stored_len = len;
                                    }
                                }
                                produced += len;
                                const uint64_t at = (src0 + len) * 8;
                                // block boundary (a stored block ends on a byte)
                                if (last) {
                                    final_seen = 1;
                                    end_byte = at >> 3;
                                    end_bits = at;
                                    state = kDone;
                                } else if (at == target) {
                                    end_byte = at >> 3;
                                    end_bits = at;
                                    state = kDone;
                                } else if (at > target) {
                                    status = ZK_DATA_ERROR;
                                    state = kDone;
                                } else {
                                    fr.seek_bits(at);
                                }
                            }
                        }
                    } else {
                        const int hrc = stored_len ? ZK_BUF_ERROR : parse_huffman_header<false>(ws, br, type, 0);
                        if (hrc != ZK_OK) {
                            status = hrc;
                            state = kDone;
                        } else {
                            // short codes into this lane's columns of the shared tables (the global tables repeat
                            // with the short period for codes that fit it; a longer code leaves its slot empty)
                            for (int i = 0; i < (1 << kLaneLitBits); ++i) {
                                const uint16_t e = ws.lit.lut[i];
                                s_lit[i * 32 + lane] = (e >> 12) <= kLaneLitBits ? e : (uint16_t)0;
                            }
                            for (int i = 0; i < (1 << kLaneDistBits); ++i) {
                                const uint16_t e = ws.dist.lut[i];
                                s_dist[i * 32 + lane] = (e >> 12) <= kLaneDistBits ? e : (uint16_t)0;
                            }
                            long_lit.load(ws.lit, kLaneLitBits + 1);
                            long_dist.load(ws.dist, kLaneDistBits + 1);
                            fr.seek_bits(br.bit_pos());
                            state = kSymbols;
                        }
                    }
                }
            } else if (state == kSymbols) {
                // ---- one token per lane.  Literals leave through a few instructions; a match or the end of a block
                // branches away (the lanes with a literal wait for them)
                fr.refill(true);
                int sym;
                {
                    const uint32_t e = s_lit[fr.peek(kLaneLitBits) * 32 + lane];
                    if (e) {
                        fr.drop((int)(e >> 12));
                        sym = (int)(e & 0x0FFFu);
                    } else {
                        int index, len;
                        long_lit.find((uint32_t)fr.hold, kLaneLitBits + 1, index, len);
                        if (len) {
                            fr.drop(len);
                            sym = ws.lit.sorted[index];
                        } else {
                            sym = decode_sym(ws.lit, fr);
                        }
                    }
                }
                if (sym < 256) {
                    if (kForeign && fr.bit_pos() > target && !extend()) {  // first: what lies behind the target is not this chunk's
                        status = kLaneOverrun;
                        state = kDone;
                    } else if (sym < 0) {
                        status = ZK_DATA_ERROR;
                        state = kDone;
                    } else if (produced >= kOutCap) {
                        status = ZK_BUF_ERROR;
                        state = kDone;
                    } else {
                        sink.put((uint32_t)sym, 1u);
                        ++produced;
                    }
                } else if (sym == 256) {
                    const uint64_t bits = fr.bit_pos();
                    if (kForeign && bits > target && !extend()) {  // only the end-of-block code lies in the next chunk's range
                        status = kLaneOverrun;
                        state = kDone;
                    } else if (last) {
                        final_seen = 1;
                        end_byte = (bits + 7) >> 3;
                        end_bits = bits;
                        state = kDone;
                    } else if (bits == target) {
                        end_byte = bits >> 3;
                        end_bits = bits;
                        state = kDone;
                    } else if (bits > target) {
                        status = kForeign ? kLaneOverrun : ZK_DATA_ERROR;
                        state = kDone;
                    } else {
                        state = kHeader;
                    }
                } else {
                    // length extra (5) + distance code (15) + distance extra (13) = 33 bits: one more refill
                    fr.refill(true);
                    // RFC 1951 3.2.5 in arithmetic (a table indexed per lane would serialise in the constant cache)
                    const int lcode = sym - 257;
                    const int le = (lcode < 8 || lcode >= 28) ? 0 : (lcode >> 2) - 1;
                    const int lbase = lcode < 8 ? 3 + lcode : lcode >= 28 ? 258 : 3 + ((4 + (lcode & 3)) << le);
                    const uint32_t length = (uint32_t)lbase + fr.peek(le);
                    fr.drop(le);
                    int dsym;
                    const uint32_t e = s_dist[fr.peek(kLaneDistBits) * 32 + lane];
                    if (e) {
                        fr.drop((int)(e >> 12));
                        dsym = (int)(e & 0x0FFFu);
                    } else {
                        int index, len;
                        long_dist.find((uint32_t)fr.hold, kLaneDistBits + 1, index, len);
                        if (len) {
                            fr.drop(len);
                            dsym = ws.dist.sorted[index];
                        } else {
                            dsym = decode_sym(ws.dist, fr);
                        }
                    }
                    const int dcl = min(max(dsym, 0), 29);
                    const int de = dcl < 4 ? 0 : (dcl >> 1) - 1;
                    const uint32_t dbase = dcl < 4 ? 1u + (uint32_t)dcl : 1u + ((2u + (uint32_t)(dcl & 1)) << de);
                    const uint32_t dist = dbase + fr.peek(de);
                    fr.drop(de);
                    if (kForeign && fr.bit_pos() > target && !extend()) {
                        status = kLaneOverrun;
                        state = kDone;
                    } else if (lcode > 28 || dsym < 0 || dsym > 29 || dist > (uint32_t)kWindow || dist > produced + reach_before) {
                        status = ZK_DATA_ERROR;
                        state = kDone;
                    } else if (produced + length > kOutCap) {
                        status = ZK_BUF_ERROR;
                        state = kDone;
                    } else {
                        sink.put((0x8000u | length) | ((dist - 1u) << 16), 2u);
                        produced += length;
                    }
                }
                // ---- up to two more literals while this lane's next code is a short literal code and its bits are
                // already in `hold`: no refill, no votes, no state changes.  (Measured at 1 GiB mixed: 0 / 1 / 2 / 3 / 4
                // extra literals per trip -> 8.1 / 6.3 / 5.9 / 6.5 / 7.0 ms: every attempt that fails -- the next token is
                // a match -- costs its dozen instructions, and the longest blocks alternate literals and matches.)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const uint32_t e = s_lit[fr.peek(kLaneLitBits) * 32 + lane];
                    const bool lit = state == kSymbols && e != 0u && (e & 0x0F00u) == 0u && fr.held >= kLaneLitBits &&
                                     produced < kOutCap && (!kForeign || fr.bit_pos() + (e >> 12) <= target);
                    if (lit) {
                        fr.drop((int)(e >> 12));
                        sink.put(e & 0xFFu, 1u);
                        ++produced;
                    }
                }
            }
        }
        const uint32_t units_tail = sink.finish(out);
        const uint32_t units = extended ? units_a + units_tail : units_tail;
        // a block that used bits past the end of the stream is not a block of this path
        if (live && status == ZK_OK && end_byte > m) status = ZK_BUF_ERROR;
        if (live) {
            lc[k].units = units;
            lc[k].units_a = extended ? units_a : units;
            lc[k].slot_off_b = slot_b;
            lc[k].ext = extended ? 1u : 0u;
            lc[k].stored_src = stored_src;
            lc[k].stored_len = stored_len;
            lc[k].status = status;
            lc[k].out_len = produced;
            lc[k].end_byte = end_byte;
            lc[k].end_bits = end_bits;
            lc[k].final_seen = final_seen;
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
    }
}

// Verify the blocks, prefix-sum their sizes, build the coarse chunk table for the window chain.
// verdict: [0] coarse chunks, [1] 1 = this path is valid, [2] status, [3] total output, [4] end byte, [5] final seen,
// [6] 1 = a chunk failed or did not end where the next one starts, [7] chunks behind false entry points (dead)
__global__ void __launch_bounds__(1024) plan_lanes_kernel(LaneChunk *lc, uint32_t n, ChunkInfo *coarse,
                                                          uint32_t *num_coarse, uint64_t *verdict, uint64_t cap, uint64_t m,
                                                          uint32_t trailer, bool shard_body) {
    __shared__ uint64_t warp_sum[32];
    __shared__ uint64_t carry;
    __shared__ int bad;
    __shared__ uint32_t dead_count, last_alive;
    if (threadIdx.x == 0) {
        carry = 0;
        bad = 0;
        dead_count = 0;
        last_alive = n - 1;
    }
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        uint64_t v = 0;
        if (i < n) {
            // A chunk that went on through the next chunk's range (ext: the entry point there was false) makes that chunk
            // dead; a dead chunk's own ext flag means nothing.  So in a run of ext flags the chunks are alive, dead,
            // alive, ...: chunk i is alive when an even number of ext flags stand directly in front of it.
            uint32_t run = 0;
            while (run < i && lc[i - 1 - run].ext) ++run;
            const bool alive = (run & 1u) == 0;
            if (!alive) {
                lc[i].out_len = 0;
                lc[i].units = 0;
                lc[i].units_a = 0;
                lc[i].stored_len = 0;
                lc[i].dead = 1;  // (ext stays as the lane left it: other threads count the flags)
                atomicAdd(&dead_count, 1u);
            } else {
                lc[i].dead = 0;
                v = lc[i].out_len;
                const bool ext = lc[i].ext != 0 && i + 1 < n;
                const uint64_t must_end = ext ? lc[i + 1].target_bits : lc[i].target_bits;
                const bool last = ext ? i + 2 == n : i + 1 == n;
                if (last) last_alive = i;
                // every chunk but the last must end exactly on the next living one; the last must close the stream
                const bool last_ok = shard_body ? (!lc[i].final_seen && lc[i].end_bits == must_end) : (lc[i].final_seen != 0);
                if (lc[i].status != ZK_OK || (last ? !last_ok : (lc[i].final_seen || lc[i].end_bits != must_end))) bad = 1;
                // a stream cut inside its trailer: the warp-per-chunk decoder reports what the reference reports (BR:48-57)
                if (last && lc[i].end_byte + trailer > m) bad = 1;
            }
        }
        uint64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint64_t before = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += warp_sum[w];
        const uint64_t start = carry + before + incl - v;
        if (i < n) lc[i].out_begin = start;
        __syncthreads();
        if (threadIdx.x == 1023) carry = start + v;
        __syncthreads();
    }
    const uint64_t total = carry;
    const uint32_t nc = (n + kCoarse - 1) / kCoarse;
    for (uint32_t c = threadIdx.x; c < nc; c += 1024) {
        const uint32_t first = c * kCoarse, lastb = min(n, first + kCoarse) - 1;
        coarse[c].in_begin = lc[first].in_begin;
        coarse[c].in_end = lc[lastb].in_end;
        coarse[c].out_begin = lc[first].out_begin;
        coarse[c].out_len = lc[lastb].out_begin + lc[lastb].out_len - lc[first].out_begin;
        coarse[c].end_byte = lc[lastb].end_byte;
        coarse[c].status = ZK_OK;
        coarse[c].final_seen = lc[lastb].final_seen;
        coarse[c].retries = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        num_coarse[0] = nc;
        verdict[0] = nc;
        verdict[1] = (!bad && total <= cap) ? 1 : 0;
        verdict[2] = (uint64_t)(int64_t)ZK_OK;
        verdict[3] = total;
        verdict[4] = lc[last_alive].end_byte;
        verdict[5] = (uint64_t)lc[last_alive].final_seen;
        verdict[6] = (uint64_t)bad;
        verdict[7] = dead_count;
        verdict[8] = lc[last_alive].end_bits;
    }
}

// Expand the token stream of every block into the dense symbol array, one WARP per block.  32 units per pass: a warp
// scan of the token lengths gives every token its position, the literals of the pass are stored by their lanes, the
// matches are done in stream order with all 32 lanes copying.  A source before the block becomes 0x8000 | (index into the
// 32 KiB before the block); resolve_coarse_kernel re-bases those afterwards.  Many warps per SM hide the L2 round trip
// of each copy.
// kForeign: the chunks of a foreign stream (decode_lanes_kernel<true>) -- stored blocks inside the token stream, matches of
// any length with all their loads in flight (long runs and repeats decode to megabytes per chunk there); the plain version
// keeps the 30 registers that let 64 warps share an SM.
constexpr int kExpandWarps = 8;
template <bool kForeign>
__global__ void __launch_bounds__(kExpandWarps * 32) expand_blocks_kernel(const uint8_t *__restrict__ in,
                                                                        const LaneChunk *__restrict__ lc, uint32_t n,
                                                                        const uint16_t *__restrict__ slots, uint16_t *dense,
                                                                        const uint32_t *__restrict__ perm,
                                                                        const uint16_t *__restrict__ spill) {
    const int lane = threadIdx.x & 31;
    const uint32_t slot = blockIdx.x * kExpandWarps + (threadIdx.x >> 5);
    if (slot >= n) return;
    // longest blocks first and blocks of similar size in one CTA (the order the lane decoder used): a CTA keeps its
    // slot until its slowest warp is done
    const uint32_t j = perm[slot];
    uint16_t *out = dense + lc[j].out_begin;
    if (lc[j].stored_len) {  // a stored block: its payload is the output
        const uint8_t *src = in + lc[j].stored_src;
        for (uint32_t i = lane; i < lc[j].stored_len; i += 32) out[i] = src[i];
        return;
    }
    const uint32_t units = lc[j].units;
    const uint16_t *tok = slots + lc[j].slot_off;
    // foreign chunks: units from units_a on are in the spill pool (the chunk went on through a false entry point's range)
    const uint32_t units_a = kForeign ? lc[j].units_a : units;
    const uint16_t *tok_b = kForeign ? spill + lc[j].slot_off_b : tok;
    auto unit = [&](uint32_t x) -> uint32_t { return (!kForeign || x < units_a) ? tok[x] : tok_b[x - units_a]; };
    uint32_t base_out = 0;
    uint32_t carry_prev = 0;  // last unit of the previous pass
    for (uint32_t u0 = 0; u0 < units; u0 += 32) {
        const uint32_t u = u0 + lane;
        const uint32_t me = u < units ? unit(u) : 0u;
        const uint32_t nxt = u + 1 < units ? unit(u + 1) : 0u;
        uint32_t prevu = __shfl_up_sync(0xFFFFFFFFu, me, 1);
        if (lane == 0) prevu = carry_prev;
        carry_prev = __shfl_sync(0xFFFFFFFFu, me, 31);
        const bool is_dist = (prevu & 0x8000u) != 0 && u < units;
        const bool is_match = !is_dist && (me & 0x8000u) != 0 && u < units;
        const bool is_lit = !is_dist && !is_match && u < units;
        uint32_t len = is_match ? (me & 0x7FFFu) : (is_lit ? 1u : 0u);
        // a match head of length 0 is a stored block of a foreign stream (decode_lanes_kernel<true>): the four units
        // behind it hold the byte offset of its payload and its length
        uint32_t stored_at = 0;
        const bool is_stored = kForeign && is_match && len == 0u;
        if (is_stored) {
            const uint32_t a = unit(u + 1) & 0x7FFFu, b = unit(u + 2) & 0x7FFFu, cc = unit(u + 3) & 0x7FFFu, d = unit(u + 4);
            stored_at = a | (b << 15) | ((cc & 3u) << 30);
            len = (cc >> 2) | (d << 13);
        }
        uint32_t incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        const uint32_t off = base_out + incl - len;
        if (is_lit) out[off] = (uint16_t)me;
        uint32_t todo = __ballot_sync(0xFFFFFFFFu, is_match);
        __syncwarp();
        while (todo) {
            const int src_lane = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t moff = __shfl_sync(0xFFFFFFFFu, off, src_lane);
            const uint32_t mlen = __shfl_sync(0xFFFFFFFFu, len, src_lane);
            const uint32_t mdist = __shfl_sync(0xFFFFFFFFu, is_stored ? 0xFFFFFFFFu : nxt, src_lane) + 1u;  // 0: a stored block
            if (kForeign && mdist == 0u) {
                // up to 64 KiB of payload: eight loads per lane in flight
                const uint32_t msrc = __shfl_sync(0xFFFFFFFFu, stored_at, src_lane);
                for (uint32_t i0 = 0; i0 < mlen; i0 += 256) {
                    uint8_t v[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint32_t i = i0 + 32 * q + lane;
                        v[q] = i < mlen ? __ldg(in + msrc + i) : (uint8_t)0;
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint32_t i = i0 + 32 * q + lane;
                        if (i < mlen) out[moff + i] = v[q];
                    }
                }
            } else if (mdist >= mlen && mlen <= 32u) {  // the common case: one pass, no overlap
                if ((uint32_t)lane < mlen) {
                    const int32_t rel = (int32_t)moff - (int32_t)mdist + lane;  // position inside this block
                    out[moff + lane] = rel >= 0 ? __ldcg(out + rel) : (uint16_t)(0x8000u | (uint32_t)(kWindow + rel));
                }
            } else if (!kForeign) {
#pragma unroll 1
                for (uint32_t i = lane; i < mlen; i += 32) {
                    const uint32_t kk = mdist >= mlen ? i : i % mdist;
                    const int32_t rel = (int32_t)moff - (int32_t)mdist + (int32_t)kk;
                    const uint16_t v = rel >= 0 ? __ldcg(out + rel) : (uint16_t)(0x8000u | (uint32_t)(kWindow + rel));
                    out[moff + i] = v;
                }
            } else {
                // a long or overlapping match: every source lies before the match (i % dist < dist), so all its loads
                // are issued before the first store
                const bool overlap = mdist < mlen;
                uint16_t v[9];  // 258 symbols at most
#pragma unroll
                for (int q = 0; q < 9; ++q) {
                    const uint32_t i = 32 * q + lane;
                    if (32u * q >= mlen) break;
                    const uint32_t kk = overlap ? i % mdist : i;
                    const int32_t rel = (int32_t)moff - (int32_t)mdist + (int32_t)kk;
                    v[q] = i >= mlen ? (uint16_t)0 : rel >= 0 ? __ldcg(out + rel) : (uint16_t)(0x8000u | (uint32_t)(kWindow + rel));
                }
#pragma unroll
                for (int q = 0; q < 9; ++q) {
                    const uint32_t i = 32 * q + lane;
                    if (i < mlen) out[moff + i] = v[q];
                }
            }
            __syncwarp();
        }
        base_out += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
}

// Re-base the symbolic references of one coarse chunk in place, block by block: a reference into an earlier block of the
// same coarse chunk becomes that block's (already final) symbol, the others are re-expressed relative to the 32 KiB before
// the coarse chunk, which is what the window chain resolves.
__global__ void __launch_bounds__(1024) resolve_coarse_kernel(const LaneChunk *__restrict__ lc, uint32_t n, uint16_t *dense) {
    const uint32_t first = blockIdx.x * kCoarse, end = min(n, first + kCoarse);
    const int64_t coarse_begin = (int64_t)lc[first].out_begin;
    for (uint32_t j = first + 1; j < end; ++j) {
        const int64_t ob = (int64_t)lc[j].out_begin;
        const uint32_t len = lc[j].out_len;
        // eight symbols per thread and step (one 16-byte load; most vectors hold no reference and are left alone);
        // the unaligned head and the tail go one by one
        const uint32_t head = tmin<uint32_t>(len, (uint32_t)((8 - (ob & 7)) & 7));
        const uint32_t nvec = (len - head) >> 3;
        for (uint32_t v = threadIdx.x; v < nvec; v += 1024) {
            uint16_t *p = dense + ob + head + (uint64_t)v * 8;
            uint4 w = *reinterpret_cast<const uint4 *>(p);
            if (((w.x | w.y | w.z | w.w) & 0x80008000u) == 0) continue;
            uint32_t ws[4] = {w.x, w.y, w.z, w.w};
            uint16_t got[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {  // the loads first ...
                const uint32_t s16 = (ws[e >> 1] >> (16 * (e & 1))) & 0xFFFFu;
                const int64_t pos = ob - kWindow + (int64_t)(s16 & 0x7FFFu);
                got[e] = ((s16 & 0x8000u) && pos >= coarse_begin) ? __ldcg(dense + pos) : (uint16_t)0;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {  // ... then the symbols
                const uint32_t s16 = (ws[e >> 1] >> (16 * (e & 1))) & 0xFFFFu;
                if (s16 & 0x8000u) {
                    const int64_t pos = ob - kWindow + (int64_t)(s16 & 0x7FFFu);
                    const uint32_t r = pos >= coarse_begin ? (uint32_t)got[e] : (0x8000u | (uint32_t)(pos - (coarse_begin - kWindow)));
                    ws[e >> 1] = (ws[e >> 1] & ~(0xFFFFu << (16 * (e & 1)))) | (r << (16 * (e & 1)));
                }
            }
            *reinterpret_cast<uint4 *>(p) = make_uint4(ws[0], ws[1], ws[2], ws[3]);
        }
        for (uint32_t i = threadIdx.x; i < head + ((len - head) & 7u); i += 1024) {
            const uint32_t at = i < head ? i : head + nvec * 8 + (i - head);
            const uint16_t s16 = dense[ob + at];
            if (s16 & 0x8000u) {
                const int64_t pos = ob - kWindow + (int64_t)(s16 & 0x7FFFu);
                dense[ob + at] = pos >= coarse_begin ? __ldcg(dense + pos)
                                                     : (uint16_t)(0x8000u | (uint32_t)(pos - (coarse_begin - kWindow)));
            }
        }
        __syncthreads();  // the next block may refer to what was just written
    }
}

}  // namespace

int run_inflate(zk_ctx *c, const InflateJob &job) {
    cudaStream_t st = c->stream;
    const uint64_t m = job.m;
    // ---- scratch ------------------------------------------------------------------------------------
    uint32_t span_shift = 9;  // >= 512-byte spans, at most ~4096 spans
    while ((m >> span_shift) > 4096) ++span_shift;
    const uint32_t num_spans = (uint32_t)(m >> span_shift) + 1;
    // lane-per-block path: blocks the scratch is reserved for (this engine's blocks decode to 32 KiB; a block needs
    // at least ~16 bytes of stream); below a few hundred blocks the warp-per-chunk path is used anyway
    uint64_t lane_blocks_reserved = tmin<uint64_t>(job.cap / 16384 + 1024, m / 16 + 1);
    if (lane_blocks_reserved < 512) lane_blocks_reserved = 0;
    const uint32_t pieces_reserved = (uint32_t)((m + kPieceBytes - 1) / kPieceBytes);
    // foreign streams: one candidate chunk per 8 KiB span (block_find_kernel), a unit of token space per bit of stream
    // (up to 2 GiB of stream in one call: the path's scratch -- 16 B of token space per compressed byte above all -- is reserved
    //  with everything else before the stream has been looked at; larger streams without markers go through the stream handle,
    //  which decodes them in pieces)
    const bool foreign_ok = c->opt_foreign_path && m >= kForeignMinBytes && m < (1ull << 31);
    const uint64_t find_spans = foreign_ok ? (m + kFindSpan - 1) / kFindSpan : 0;
    const uint32_t max_chunks = (uint32_t)tmax<uint64_t>(tmax<uint64_t>(num_spans + 1, lane_blocks_reserved / kCoarse + 2),
                                                         find_spans * kFindSub / kCoarse + 2);
    c->arena.reset();
    if (job.container != 3) c->shard.valid = false;
    const size_t need = 4096 + (size_t)num_spans * 4 + (size_t)max_chunks * sizeof(ChunkInfo) + 1024 +
                        (size_t)(max_chunks + 1) * kWindow * 3 + (job.cap + 64) * 2 + adler_partials_bytes(job.cap) + crc_partials_bytes(job.cap) + 8192 +
                        (size_t)pieces_reserved * 12 + 1024 +
                        (size_t)lane_blocks_reserved * (12 + sizeof(LaneChunk) + sizeof(WarpState) + (size_t)kSlotSyms * 2 + 1024) +
                        (size_t)find_spans * kFindSub * (24 + sizeof(LaneChunk) + sizeof(WarpState) + 8) + (foreign_ok ? (size_t)(m * 8 + 256) * 2 + (size_t)(m + (4ull << 20)) * 2 + 16384 : 0);
    if (c->arena.reserve(need) != ZK_OK) return ZK_MEM_ERROR;
    int32_t *d_header = static_cast<int32_t *>(c->arena.take(256));
    uint32_t *d_cand = static_cast<uint32_t *>(c->arena.take((size_t)num_spans * 4));
    ChunkInfo *d_chunks = static_cast<ChunkInfo *>(c->arena.take((size_t)max_chunks * sizeof(ChunkInfo)));
    uint32_t *d_num = static_cast<uint32_t *>(c->arena.take(256));
    uint64_t *d_verdict = static_cast<uint64_t *>(c->arena.take(256));
    uint32_t *d_flag = static_cast<uint32_t *>(c->arena.take(256));
    uint32_t *d_adler = static_cast<uint32_t *>(c->arena.take(256));
    void *d_partials = c->arena.take(tmax(adler_partials_bytes(job.cap), crc_partials_bytes(job.cap)));
    // windows[-1] is the window before the stream: the preset dictionary right-aligned, zeros otherwise
    uint8_t *d_windows = static_cast<uint8_t *>(c->arena.take((size_t)(max_chunks + 1) * kWindow)) + kWindow;
    uint16_t *d_maps = static_cast<uint16_t *>(c->arena.take((size_t)max_chunks * kWindow * 2));
    uint16_t *d_sym = static_cast<uint16_t *>(c->arena.take((job.cap + 64) * 2));

    // ---- header -------------------------------------------------------------------------------------
    int32_t hdr[2] = {ZK_OK, (int32_t)job.body_begin};
    if (job.container == 0) {
        header_kernel<<<1, 32, 0, st>>>(job.d_in, m, d_header, job.d_dict != nullptr, job.dict_id);
        c->launches++;
        ZK_CUDA(cudaMemcpyAsync(hdr, d_header, 8, cudaMemcpyDeviceToHost, st));
        ZK_CUDA(cudaStreamSynchronize(st));
    }
    const int32_t hst = hdr[0];
    const uint32_t body_begin = (uint32_t)hdr[1];
    // a shard body may reach hist_len bytes back into the output of the shards before it: the same mechanism as a preset
    // dictionary, except that the window's CONTENT arrives later (run_inflate_shard_finish)
    // (a raw deflate stream may be given a dictionary too: the streaming inflate decodes the end of a stream that way,
    // with the last 32 KiB of what it has produced so far as the dictionary)
    const uint32_t dict_len = job.container == 3 ? tmin<uint32_t>(job.dict_len, (uint32_t)kWindow)
                              : (job.d_dict && (body_begin == 6 || job.container == 1)) ? tmin<uint32_t>(job.dict_len, (uint32_t)kWindow) : 0u;
    ZK_CUDA(cudaMemsetAsync(d_windows - kWindow, 0, kWindow, st));
    if (dict_len && job.container != 3)
        ZK_CUDA(cudaMemcpyAsync(d_windows - dict_len, job.d_dict + (job.dict_len - dict_len), dict_len, cudaMemcpyDeviceToDevice, st));
    if (hst != ZK_OK) {
        // (hdr, 0): nothing decoded; FDICT consumes the 4 DICTID bytes (IS:41)
        set_result_kernel<<<1, 32, 0, st>>>(job.d_result, 0, hst == ZK_NEED_DICT ? 6 : (m < 2 ? m : 2), hst, 1);
        c->launches++;
        ZK_CUDA(cudaGetLastError());
        return ZK_OK;
    }
    ZK_CUDA(cudaMemsetAsync(d_flag, 0, 4, st));
    // ---- lane-per-block path: every marker is a block start -------------------------------------------------
    bool lane_path_done = false;
    uint64_t verdict[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    {
        const uint32_t pieces = (uint32_t)((m + kPieceBytes - 1) / kPieceBytes);
        uint32_t *d_counts = static_cast<uint32_t *>(c->arena.take((size_t)pieces * 4 + 16));
        uint64_t *d_piece_off = static_cast<uint64_t *>(c->arena.take((size_t)pieces * 8 + 16));
        uint64_t *d_total = static_cast<uint64_t *>(c->arena.take(256));
        if (d_counts && d_piece_off && d_total) {
            marker_compact_kernel<false><<<pieces, 256, 0, st>>>(job.d_in, m, d_counts, nullptr, nullptr, 0);
            marker_scan_kernel<<<1, 1024, 0, st>>>(d_counts, pieces, d_piece_off, d_total);
            c->launches += 2;
            uint64_t num_markers = 0;
            ZK_CUDA(cudaMemcpyAsync(&num_markers, d_total, 8, cudaMemcpyDeviceToHost, st));
            ZK_CUDA(cudaStreamSynchronize(st));
            const uint64_t n = num_markers + 1;
            // worth it from a few hundred blocks on, only if the scratch was reserved for it, and only when the markers are
            // as close as this engine's blocks are (32 KiB of output each): pigz's 128 KiB pieces would fail on the slot size
            // after decoding a third of the stream
            if (num_markers >= 256 && n <= lane_blocks_reserved && m <= n * (uint64_t)(kSlotSyms + 64) && !job.prefix) {
                uint64_t *d_pos = static_cast<uint64_t *>(c->arena.take((size_t)n * 8));
                LaneChunk *d_lc = static_cast<LaneChunk *>(c->arena.take((size_t)n * sizeof(LaneChunk)));
                WarpState *d_tables = static_cast<WarpState *>(c->arena.take((size_t)n * sizeof(WarpState)));
                uint16_t *d_slots = static_cast<uint16_t *>(c->arena.take((size_t)n * kSlotSyms * 2));
                if (d_pos && d_lc && d_tables && d_slots) {
                    marker_compact_kernel<true><<<pieces, 256, 0, st>>>(job.d_in, m, nullptr, d_piece_off, d_pos, num_markers);
                    lane_table_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, st>>>(d_pos, (uint32_t)num_markers, m, d_lc, body_begin, job.begin_bit);
                    uint32_t *d_perm = static_cast<uint32_t *>(c->arena.take((size_t)n * 4));
                    lane_order_kernel<<<1, 1024, 0, st>>>(d_lc, (uint32_t)n, d_perm);
                    c->launches += 1;
                    {
                        ScopedTimer t(c, 4);
                        const int lane_smem = kLaneWarps * kLaneWarpSmem;
                        const uint32_t tasks = (uint32_t)((n + 31) / 32);
                        // a warp per SM before a second one -- unless several calls are to run side by side (the slabs of the
                        // pipelined host-buffer inflate): then whole CTAs, so that a call occupies few SMs
                        const uint32_t grid = c->opt_lane_pack ? tmin<uint32_t>((tasks + kLaneWarps - 1) / kLaneWarps, (uint32_t)c->sm_count)
                                                               : tmin<uint32_t>(tasks, (uint32_t)c->sm_count);
                        ZK_CUDA(cudaFuncSetAttribute(decode_lanes_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lane_smem));
                        ZK_CUDA(cudaMemsetAsync(d_total, 0, 4, st));  // reused as the task counter
                        decode_lanes_kernel<false><<<grid, kLaneWarps * 32, lane_smem, st>>>(
                            job.d_in, m, d_lc, (uint32_t)n, d_tables, d_slots, reinterpret_cast<uint32_t *>(d_total), d_perm, dict_len,
                            nullptr, nullptr, 0);
                    }
                    plan_lanes_kernel<<<1, 1024, 0, st>>>(d_lc, (uint32_t)n, d_chunks, d_num, d_verdict, job.cap, m,
                                                          job.container == 0 ? 4u : job.container == 2 ? 8u : 0u, job.container == 3);
                    c->launches += 4;
                    ZK_CUDA(cudaMemcpyAsync(verdict, d_verdict, sizeof(verdict), cudaMemcpyDeviceToHost, st));
                    ZK_CUDA(cudaStreamSynchronize(st));
                    if (verdict[1] == 1 && verdict[0] <= max_chunks) {
                        ScopedTimer t(c, 5);
                        expand_blocks_kernel<false><<<(uint32_t)((n + kExpandWarps - 1) / kExpandWarps), kExpandWarps * 32, 0, st>>>(
                            job.d_in, d_lc, (uint32_t)n, d_slots, d_sym, d_perm, nullptr);
                        resolve_coarse_kernel<<<(uint32_t)verdict[0], 1024, 0, st>>>(d_lc, (uint32_t)n, d_sym);
                        c->launches += 2;
                        lane_path_done = true;
                    }
                }
            }
        }
    }

    // ---- a stream somebody else made: entry points at the dynamic block headers block_find_kernel accepts ------------------
    const uint64_t first_bits = (uint64_t)body_begin * 8 + job.begin_bit;
    if (!lane_path_done && foreign_ok) {
        const uint32_t find_ctas = (uint32_t)find_spans;
        const uint32_t spans = find_ctas * kFindSub;  // candidate slots
        uint64_t *d_cand64 = static_cast<uint64_t *>(c->arena.take((size_t)spans * 8));
        uint32_t *d_flags = static_cast<uint32_t *>(c->arena.take((size_t)spans * 4 + 16));
        uint64_t *d_off = static_cast<uint64_t *>(c->arena.take((size_t)spans * 8 + 16));
        uint64_t *d_count = static_cast<uint64_t *>(c->arena.take(256));
        LaneChunk *d_lc = static_cast<LaneChunk *>(c->arena.take((size_t)(spans + 1) * sizeof(LaneChunk)));
        WarpState *d_tables = static_cast<WarpState *>(c->arena.take((size_t)(spans + 1) * sizeof(WarpState)));
        uint32_t *d_perm = static_cast<uint32_t *>(c->arena.take((size_t)(spans + 1) * 4));
        uint32_t *d_redo = static_cast<uint32_t *>(c->arena.take((size_t)(spans + 1) * 4));
        uint16_t *d_slots = static_cast<uint16_t *>(c->arena.take((size_t)(m * 8 + 256) * 2));
        // units of chunks that go on through a false entry point's range: a unit per bit of that range, rarely needed
        const uint64_t spill_units = m + (4ull << 20);
        uint16_t *d_spill = static_cast<uint16_t *>(c->arena.take((size_t)spill_units * 2));
        unsigned long long *d_spill_next = static_cast<unsigned long long *>(c->arena.take(256));
        if (d_cand64 && d_flags && d_off && d_count && d_lc && d_tables && d_perm && d_redo && d_slots && d_spill && d_spill_next) {
            uint64_t found = 0;
            {
                ScopedTimer t(c, 5);
                ZK_CUDA(cudaFuncSetAttribute(block_find_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FindSmem)));
                block_find_kernel<<<find_ctas, kFindThreads, sizeof(FindSmem), st>>>(job.d_in, m, first_bits, d_cand64, d_flags);
                marker_scan_kernel<<<1, 1024, 0, st>>>(d_flags, spans, d_off, d_count);
                c->launches += 2;
            }
            ZK_CUDA(cudaMemcpyAsync(&found, d_count, 8, cudaMemcpyDeviceToHost, st));
            ZK_CUDA(cudaStreamSynchronize(st));
            const bool dbg = getenv("ZK_FOREIGN_DEBUG") != nullptr;
            if (dbg) fprintf(stderr, "[zk foreign] m %llu candidate slots %u entry points %llu\n", (unsigned long long)m, spans, (unsigned long long)found);
            if (found >= kForeignMinChunks + (job.prefix ? 1 : 0)) {
                uint64_t n = found + 1;
                foreign_table_kernel<<<(spans + 256) / 256, 256, 0, st>>>(d_cand64, d_flags, d_off, spans, found, m, first_bits, d_lc);
                // prefix mode: the last chunk may be cut; it stays in the table as the target of the one before it and is
                // left alone otherwise
                if (job.prefix) n -= 1;
                lane_order_kernel<<<1, 1024, 0, st>>>(d_lc, (uint32_t)n, d_perm);
                c->launches += 2;
                const int lane_smem = kLaneWarps * kLaneWarpSmem;
                ZK_CUDA(cudaFuncSetAttribute(decode_lanes_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lane_smem));
                const uint32_t *d_todo = d_perm;
                uint64_t todo = n;
                ZK_CUDA(cudaMemsetAsync(d_spill_next, 0, 8, st));
                std::vector<LaneChunk> host_lc;
                std::vector<uint32_t> redo;
                for (int round = 0; round < kForeignRounds && !lane_path_done; ++round) {
                    {
                        ScopedTimer t(c, 4);
                        const uint32_t tasks = (uint32_t)((todo + 31) / 32);
                        ZK_CUDA(cudaMemsetAsync(d_count, 0, 4, st));  // the task counter
                        decode_lanes_kernel<true><<<tmin<uint32_t>(tasks, (uint32_t)c->sm_count), kLaneWarps * 32, lane_smem, st>>>(
                            job.d_in, m, d_lc, (uint32_t)todo, d_tables, d_slots, reinterpret_cast<uint32_t *>(d_count), d_todo, dict_len,
                            d_spill, d_spill_next, spill_units);
                    }
                    plan_lanes_kernel<<<1, 1024, 0, st>>>(d_lc, (uint32_t)n, d_chunks, d_num, d_verdict, job.cap, m,
                                                          job.container == 0 ? 4u : job.container == 2 ? 8u : 0u, job.prefix || job.container == 3);
                    c->launches += 2;
                    ZK_CUDA(cudaMemcpyAsync(verdict, d_verdict, sizeof(verdict), cudaMemcpyDeviceToHost, st));
                    ZK_CUDA(cudaStreamSynchronize(st));
                    if (dbg)
                        fprintf(stderr, "[zk foreign] round %d: chunks %llu decoded %llu bad %llu valid %llu total %llu coarse %llu (max %u)\n", round,
                                (unsigned long long)n, (unsigned long long)todo, (unsigned long long)verdict[6], (unsigned long long)verdict[1],
                                (unsigned long long)verdict[3], (unsigned long long)verdict[0], max_chunks);
                    if (!verdict[6]) {
                        if (verdict[1] == 1 && verdict[0] <= max_chunks) {
                            lane_path_done = true;
                        } else if (verdict[3] > job.cap) {  // the output does not fit: nothing is written
                            set_result_kernel<<<1, 32, 0, st>>>(job.d_result, 0, 0, ZK_BUF_ERROR, 1);
                            c->launches++;
                            ZK_CUDA(cudaGetLastError());
                            return ZK_OK;
                        }
                        break;
                    }
                    // a chunk that ran past its target had a false entry point behind it: merge the two and decode the merged
                    // chunk again; chunks behind the stream's last block go; anything else is the slow path's to report
                    host_lc.resize(n);
                    ZK_CUDA(cudaMemcpyAsync(host_lc.data(), d_lc, n * sizeof(LaneChunk), cudaMemcpyDeviceToHost, st));
                    ZK_CUDA(cudaStreamSynchronize(st));
                    redo.clear();
                    uint64_t w = 0;
                    bool give_up = false, incomplete = false;
                    for (uint64_t i = 0; i < n && !give_up; ++i) {
                        LaneChunk cur = host_lc[i];
                        if (cur.dead) {  // marked by the plan of this round: the picture may change, decode it again
                            cur.dead = 0;
                            cur.status = ZK_OK;
                            redo.push_back((uint32_t)w);
                            host_lc[w++] = cur;
                            continue;
                        }
                        if (cur.status == kLaneOverrun && i + 1 < n) {
                            cur.target_bits = host_lc[i + 1].target_bits;
                            cur.in_end = host_lc[i + 1].in_end;
                            cur.status = ZK_OK;
                            redo.push_back((uint32_t)w);
                            ++i;
                        } else if (cur.status != ZK_OK) {
                            // the last chunk ran out of input and everything before it is in order
                            if (job.quick_incomplete && redo.empty() && (i + 1 == n || (cur.ext && i + 2 == n)) &&
                                (cur.status == ZK_BUF_ERROR || cur.status == kLaneOverrun))
                                incomplete = true;
                            if (dbg)
                                fprintf(stderr, "[zk foreign] chunk %llu at bit %llu (target %llu): status %d after %u bytes\n", (unsigned long long)i,
                                        (unsigned long long)cur.begin_bits, (unsigned long long)cur.target_bits, cur.status, cur.out_len);
                            give_up = true;
                        }
                        host_lc[w++] = cur;
                        if (cur.status == ZK_OK && cur.final_seen && redo.empty()) break;  // the stream ends here
                    }
                    // (a last chunk that ends on the last byte without a final block is an unfinished stream too)
                    if (job.quick_incomplete && !give_up && redo.empty() && n && host_lc[w - 1].status == ZK_OK && !host_lc[w - 1].final_seen)
                        incomplete = true;
                    if (incomplete) {
                        c->quick_incomplete++;
                        // (consumed = m: the input ran out -- an output buffer that is too small reports 0 consumed)
                        set_result_kernel<<<1, 32, 0, st>>>(job.d_result, 0, m, ZK_BUF_ERROR, 1);
                        c->launches++;
                        ZK_CUDA(cudaGetLastError());
                        return ZK_OK;
                    }
                    if (give_up || redo.empty()) break;
                    n = w;
                    todo = redo.size();
                    ZK_CUDA(cudaMemcpyAsync(d_lc, host_lc.data(), n * sizeof(LaneChunk), cudaMemcpyHostToDevice, st));
                    ZK_CUDA(cudaMemcpyAsync(d_redo, redo.data(), todo * 4, cudaMemcpyHostToDevice, st));
                    ZK_CUDA(cudaStreamSynchronize(st));  // the vectors are reused
                    d_todo = d_redo;
                    c->foreign_merges += (int64_t)todo;
                }
                if (lane_path_done) {
                    ScopedTimer t(c, 5);
                    lane_order_kernel<<<1, 1024, 0, st>>>(d_lc, (uint32_t)n, d_perm);
                    expand_blocks_kernel<true><<<(uint32_t)((n + kExpandWarps - 1) / kExpandWarps), kExpandWarps * 32, 0, st>>>(
                        job.d_in, d_lc, (uint32_t)n, d_slots, d_sym, d_perm, d_spill);
                    resolve_coarse_kernel<<<(uint32_t)verdict[0], 1024, 0, st>>>(d_lc, (uint32_t)n, d_sym);
                    c->launches += 3;
                    c->foreign_chunks += (int64_t)(n - verdict[7]);
                    c->foreign_merges += (int64_t)verdict[7];
                }
            }
        }
    }
    uint64_t total = verdict[3];
    if (job.prefix && !lane_path_done) {  // no progress possible: the caller gathers more input
        set_result_kernel<<<1, 32, 0, st>>>(job.d_result, 0, 0, ZK_BUF_ERROR, 1);
        c->launches++;
        ZK_CUDA(cudaGetLastError());
        return ZK_OK;
    }
    if (!lane_path_done && m >= (1ull << 32)) {
        // the chunk path below keeps candidate positions in 32 bits: a stream of 4 GiB or more that is not cut
        // by sync markers would have to be decoded by one warp anyway -- refuse it instead of wrapping positions
        set_last_error("inflate: compressed input of 4 GiB or more without sync markers is not supported", cudaErrorInvalidValue);
        return ZK_STREAM_ERROR;
    }
    if (!lane_path_done) {
        // ---- candidates -> chunks -------------------------------------------------------------------------
        fill_u32_kernel<<<32, 256, 0, st>>>(d_cand, num_spans, kNoCand);
        {
            const uint32_t grid = (uint32_t)tmin<uint64_t>((m + 255) / 256, (uint64_t)c->sm_count * 8);
            marker_kernel<<<grid ? grid : 1, 256, 0, st>>>(job.d_in, m, span_shift, d_cand);
        }
        chunk_list_kernel<<<1, 32, 0, st>>>(d_cand, num_spans, 1, m, d_chunks, d_num, max_chunks, body_begin);
        c->launches += 3;
        // ---- size pass, repeated while candidates turn out to be false ----------------------------------------
        for (int round = 0; round < 64; ++round) {
            const uint32_t grid = (max_chunks + kInfWarps - 1) / kInfWarps;
            {
                ScopedTimer t(c, 4);
                decode_kernel<false><<<grid, kInfWarps * 32, 0, st>>>(job.d_in, m, d_chunks, d_num, nullptr, 0, dict_len, job.begin_bit);
            }
            plan_chunks_kernel<<<1, 32, 0, st>>>(d_chunks, d_num, d_verdict);
            c->launches += 2;
            ZK_CUDA(cudaMemcpyAsync(verdict, d_verdict, sizeof(verdict), cudaMemcpyDeviceToHost, st));
            ZK_CUDA(cudaStreamSynchronize(st));
            if (!verdict[1]) break;
        }
        total = verdict[3];
        if (total > job.cap) {
            set_result_kernel<<<1, 32, 0, st>>>(job.d_result, 0, 0, ZK_BUF_ERROR, 1);
            c->launches++;
            ZK_CUDA(cudaGetLastError());
            return ZK_OK;
        }
        // ---- write pass -------------------------------------------------------------------------------------------
        {
            const uint32_t nchunks = (uint32_t)verdict[0];
            const uint32_t grid = (nchunks + kInfWarps - 1) / kInfWarps;
            ScopedTimer t(c, 4);
            decode_kernel<true><<<grid ? grid : 1, kInfWarps * 32, 0, st>>>(job.d_in, m, d_chunks, d_num, d_sym, job.cap, dict_len, job.begin_bit);
            c->launches += 1;
        }
    }
    // ---- window chain over the (coarse) chunks, symbols -> bytes ------------------------------------------------
    {
        ScopedTimer t(c, 5);
        const uint32_t nch = (uint32_t)verdict[0];
        if (nch) {
            // phase A takes `group` dependent steps of ~18 us (64 KiB maps), phase B n / group steps of ~4 us
            uint32_t group = 1;
            while (group * group * 4 < nch) group <<= 1;
            const uint32_t groups = (nch + group - 1) / group;
            ZK_CUDA(cudaFuncSetAttribute(window_compose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kWindow));
            ZK_CUDA(cudaFuncSetAttribute(window_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kWindow));
            window_compose_kernel<<<groups, 1024, 4 * kWindow, st>>>(d_chunks, nch, group, d_sym, d_maps, d_flag, dict_len);
            if (job.container == 3) {
                // phase 1 of a sharded inflate ends here: hand out the shard's total tail map, keep the rest for phase 2
                ZK_CUDA(cudaFuncSetAttribute(window_chain_maps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kWindow));
                window_chain_maps_kernel<<<1, 1024, 4 * kWindow, st>>>(nch, group, d_maps, job.d_tail_map);
                c->launches += 2;
                c->shard = {true, d_chunks, d_num, d_verdict, d_flag, d_adler, d_partials, d_windows, d_maps, d_sym, nch, group, total, dict_len, job.d_in, m};
                set_result_kernel<<<1, 32, 0, st>>>(job.d_result, total, verdict[4], (int64_t)verdict[2], 0);
                c->launches++;
                ZK_CUDA(cudaGetLastError());
                return ZK_OK;
            }
            window_chain_kernel<<<1, 1024, 2 * kWindow, st>>>(nch, group, d_maps, d_windows);
            window_apply_kernel<<<nch, 1024, 0, st>>>(nch, group, d_maps, d_windows);
            c->launches += 2;
        }
        if (job.container == 3) {  // a shard without output: the identity map
            ZK_CUDA(cudaFuncSetAttribute(window_chain_maps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kWindow));
            window_chain_maps_kernel<<<1, 1024, 4 * kWindow, st>>>(0, 1, d_maps, job.d_tail_map);
            c->shard = {true, d_chunks, d_num, d_verdict, d_flag, d_adler, d_partials, d_windows, d_maps, d_sym, 0, 1, total, dict_len, job.d_in, m};
            set_result_kernel<<<1, 32, 0, st>>>(job.d_result, total, verdict[4], (int64_t)verdict[2], 0);
            c->launches += 2;
            ZK_CUDA(cudaGetLastError());
            return ZK_OK;
        }
        const uint32_t tiles = (uint32_t)((total + 16383) / 16384);
        if (tiles) materialize_kernel<<<tiles, 256, 0, st>>>(d_chunks, d_num, d_sym, d_windows, job.d_out, total, d_flag, dict_len);
        c->launches += 2;
    }
    int rc = job.container == 2 ? launch_crc32(c, 0u, job.d_out, total, d_partials, d_adler)
                                : launch_adler32(c, 1u, job.d_out, total, d_partials, d_adler);
    if (rc != ZK_OK) return rc;
    finish_inflate_kernel<<<1, 32, 0, st>>>(job.d_in, m, d_verdict, d_adler, d_flag, job.d_result, job.prefix ? 4 : job.container);
    c->launches++;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}

// Phase 2 of a sharded inflate: the window in front of the shard is known now (the tail maps of the shards before it,
// composed): resolve the shard's windows, turn its symbols into bytes, checksum them.  Uses the scratch phase 1 left in
// the context's arena: no other call on this context may come between the two.
int run_inflate_shard_finish(zk_ctx *c, const uint8_t *d_entry_window, uint8_t *d_out, uint64_t *d_result) {
    if (!c->shard.valid) return ZK_STREAM_ERROR;
    const zk_ctx::ShardState sh = c->shard;
    c->shard.valid = false;
    cudaStream_t st = c->stream;
    uint8_t *windows = static_cast<uint8_t *>(sh.d_windows);
    ZK_CUDA(cudaMemcpyAsync(windows - kWindow, d_entry_window, kWindow, cudaMemcpyDeviceToDevice, st));
    {
        ScopedTimer t(c, 5);
        if (sh.nch) {
            ZK_CUDA(cudaFuncSetAttribute(window_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kWindow));
            window_chain_kernel<<<1, 1024, 2 * kWindow, st>>>(sh.nch, sh.group, static_cast<const uint16_t *>(sh.d_maps), windows);
            window_apply_kernel<<<sh.nch, 1024, 0, st>>>(sh.nch, sh.group, static_cast<const uint16_t *>(sh.d_maps), windows);
        }
        const uint32_t tiles = (uint32_t)((sh.total + 16383) / 16384);
        if (tiles)
            materialize_kernel<<<tiles, 256, 0, st>>>(static_cast<const ChunkInfo *>(sh.d_chunks), static_cast<const uint32_t *>(sh.d_num),
                                                      static_cast<const uint16_t *>(sh.d_sym), windows, d_out, sh.total,
                                                      static_cast<uint32_t *>(sh.d_flag), sh.dict_len);
        c->launches += 3;
    }
    int rc = launch_adler32(c, 1u, d_out, sh.total, sh.d_partials, static_cast<uint32_t *>(sh.d_adler));
    if (rc != ZK_OK) return rc;
    finish_inflate_kernel<<<1, 32, 0, st>>>(sh.d_in, sh.m, static_cast<const uint64_t *>(sh.d_verdict), static_cast<const uint32_t *>(sh.d_adler),
                                            static_cast<const uint32_t *>(sh.d_flag), d_result, 3);
    c->launches++;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}

int launch_window_from_maps(zk_ctx *c, const uint16_t *d_maps, uint64_t stride_u16, uint32_t count, uint8_t *d_window) {
    ZK_CUDA(cudaFuncSetAttribute(window_from_maps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kWindow));
    window_from_maps_kernel<<<1, 1024, 2 * kWindow, c->stream>>>(d_maps, stride_u16, count, d_window);
    c->launches++;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}

}  // namespace zk
