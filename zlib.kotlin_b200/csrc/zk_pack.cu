// zk_pack.cu -- K3: bit-packing, one CTA per deflate block.
//
// Replaces StreamingBitWriter (inflate/StreamingBitWriter.kt:10-49) and the emit loops of
// DeflateStream.kt:958-1035 / 98-105 / 231-252, plus the stored path (DS:331-373, 80-95) and the zlib
// framing (DS:35-74).  The reference shifts one symbol at a time into a 32-bit accumulator and hands the
// sink one byte per call; here every thread codes 8 consecutive tokens, a block-wide prefix sum of the
// code lengths gives each thread its bit offset, the bits are OR-ed into a shared-memory image of the
// block, and the image leaves the SM as aligned 128-bit stores.
//
// Every block starts on a byte boundary of the stream (the previous block ended with an empty stored
// block), so the shared image is laid out with the same 16-byte phase as its destination.
#include "zk_common.cuh"

namespace zk {

namespace {

constexpr int kPackThreads = 256;
constexpr int kTokPerThread = 8;
constexpr int kTileTokens = kPackThreads * kTokPerThread;

__constant__ uint8_t p_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};

struct PackArgs {
    const uint8_t *in;            // shard input (for stored blocks)
    uint64_t n;
    uint32_t block_bytes;
    uint32_t num_blocks;
    const uint32_t *tokens;
    const uint32_t *block_tok_begin;
    const uint32_t *block_tok_count;
    const BlockPlan *plans;
    const uint64_t *offsets;      // byte offset of each block in the body
    const uint64_t *cursor;       // [1] != 0 -> output does not fit, write nothing
    uint8_t *out;                 // body destination
    uint32_t image_words;         // shared image capacity in 32-bit words
};

__device__ __forceinline__ int length_symbol_index(int len) {  // 0..28
    const int y = len - 3;
    if (y < 8) return y;
    if (len == 258) return 28;
    const int b = 31 - __clz(y);
    return 4 * (b - 1) + ((y >> (b - 2)) & 3);
}

// bits of one token under the block's code: value (LSB-first, <= 48 bits) and width
__device__ __forceinline__ void code_token(uint32_t tok, const uint32_t *lit, const uint32_t *dst, uint64_t &bits,
                                           int &width) {
    const int len = (int)(tok >> 16);
    if (len == 0) {
        const uint32_t c = lit[tok & 0xFFu];
        bits = c & 0xFFFFu;
        width = (int)(c >> 16);
        return;
    }
    const int li = length_symbol_index(len);
    const uint32_t lc = lit[257 + li];
    const int le = p_len_extra[li];
    // base length of symbol li: 3 + li for li < 8, else from the (b, quarter) decomposition
    int lbase;
    if (li < 8) lbase = 3 + li;
    else if (li == 28) lbase = 258;
    else {
        const int b = (li >> 2) + 1;
        lbase = 3 + ((4 + (li & 3)) << (b - 2));
    }
    uint64_t v = lc & 0xFFFFu;
    int wdt = (int)(lc >> 16);
    v |= (uint64_t)(uint32_t)(len - lbase) << wdt;
    wdt += le;
    const int dm1 = (int)(tok & 0xFFFFu);
    int ds, de, dbase;
    if (dm1 < 4) {
        ds = dm1;
        de = 0;
        dbase = dm1;
    } else {
        const int b = 31 - __clz(dm1);
        const int half = (dm1 >> (b - 1)) & 1;
        ds = 2 * b + half;
        de = b - 1;
        dbase = (2 + half) << (b - 1);
    }
    const uint32_t dc = dst[ds];
    v |= (uint64_t)(dc & 0xFFFFu) << wdt;
    wdt += (int)(dc >> 16);
    v |= (uint64_t)(uint32_t)(dm1 - dbase) << wdt;
    wdt += de;
    bits = v;
    width = wdt;
}

// OR `width` bits of `bits` into the shared image at bit offset `at`
__device__ __forceinline__ void image_or(uint32_t *image, uint64_t at, uint64_t bits, int width) {
    if (width == 0) return;
    const uint32_t word = (uint32_t)(at >> 5);
    const int sh = (int)(at & 31);
    atomicOr(&image[word], (uint32_t)(bits << sh));
    const int first = 32 - sh;
    if (width > first) {
        atomicOr(&image[word + 1], (uint32_t)(bits >> first));
        if (width > first + 32) atomicOr(&image[word + 2], (uint32_t)(bits >> (first + 32)));
    }
}

__global__ void __launch_bounds__(kPackThreads) pack_kernel(const PackArgs a) {
    extern __shared__ __align__(16) uint32_t image[];
    __shared__ uint32_t lit[kNumLitLen + 2];
    __shared__ uint32_t dst[kNumDist + 2];
    __shared__ uint64_t warp_sum[kPackThreads / 32];
    __shared__ uint64_t tile_base;

    if (a.cursor[1] != 0) return;
    const uint32_t blk = blockIdx.x;
    const BlockPlan &plan = a.plans[blk];
    const uint64_t off = a.offsets[blk];
    uint8_t *dest = a.out + off;
    const uint32_t out_bytes = plan.out_bytes;
    const uint64_t raw0 = (uint64_t)blk * a.block_bytes;
    const uint32_t raw = (uint32_t)tmin<uint64_t>(a.block_bytes, a.n - raw0);

    if (plan.type == 0) {
        // stored: 00 | LEN | NLEN | bytes.  BFINAL = 0, the five pad bits are zero.
        if (threadIdx.x == 0) {
            dest[0] = 0;
            dest[1] = (uint8_t)(raw & 0xFF);
            dest[2] = (uint8_t)(raw >> 8);
            dest[3] = (uint8_t)(~raw & 0xFF);
            dest[4] = (uint8_t)((~raw >> 8) & 0xFF);
        }
        const uint8_t *src = a.in + raw0;
        uint8_t *d = dest + 5;
        // byte copy with a 16-byte fast path when source and destination share their phase
        if (((reinterpret_cast<uintptr_t>(src) ^ reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
            uint32_t headn = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15);
            if (headn > raw) headn = raw;
            if (threadIdx.x < headn) d[threadIdx.x] = src[threadIdx.x];
            const uint32_t nvec = (raw - headn) >> 4;
            const uint4 *sv = reinterpret_cast<const uint4 *>(src + headn);
            uint4 *dv = reinterpret_cast<uint4 *>(d + headn);
            for (uint32_t i = threadIdx.x; i < nvec; i += kPackThreads) dv[i] = sv[i];
            for (uint32_t i = headn + (nvec << 4) + threadIdx.x; i < raw; i += kPackThreads) d[i] = src[i];
        } else {
            for (uint32_t i = threadIdx.x; i < raw; i += kPackThreads) d[i] = src[i];
        }
        if (threadIdx.x < 5) d[raw + threadIdx.x] = threadIdx.x >= 3 ? 0xFF : 0x00;  // sync marker: 00 | 0000 | FFFF
        return;
    }

    // shared image with the destination's 16-byte phase
    const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(dest) & 15);
    const uint32_t img_bytes = lead + out_bytes;
    const uint32_t img_words = (img_bytes + 3) >> 2;
    for (uint32_t i = threadIdx.x; i < img_words + 2 && i < a.image_words; i += kPackThreads) image[i] = 0;
    for (int i = threadIdx.x; i < kNumLitLen; i += kPackThreads) lit[i] = plan.litlen_code[i];
    if (threadIdx.x < kNumDist) dst[threadIdx.x] = plan.dist_code[threadIdx.x];
    __syncthreads();

    // block header (already a bit string starting at bit 0)
    const uint32_t hbits = plan.header_bits;
    {
        uint8_t *img8 = reinterpret_cast<uint8_t *>(image);
        const uint32_t hbytes = (hbits + 7) >> 3;
        for (uint32_t i = threadIdx.x; i < hbytes; i += kPackThreads) img8[lead + i] = plan.header[i];
    }
    if (threadIdx.x == 0) tile_base = (uint64_t)lead * 8 + hbits;
    __syncthreads();

    const uint32_t tb = a.block_tok_begin[blk], tc = a.block_tok_count[blk];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t t0 = 0; t0 < tc; t0 += kTileTokens) {
        uint64_t bits[kTokPerThread];
        int width[kTokPerThread];
        uint32_t total = 0;
        const uint32_t first = t0 + threadIdx.x * kTokPerThread;
#pragma unroll
        for (int k = 0; k < kTokPerThread; ++k) {
            width[k] = 0;
            bits[k] = 0;
            if (first + k < tc) code_token(a.tokens[tb + first + k], lit, dst, bits[k], width[k]);
            total += (uint32_t)width[k];
        }
        // block-wide exclusive prefix sum of the widths
        uint32_t incl = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_sum[wid] = incl;
        __syncthreads();
        uint64_t before = 0;
        for (int w = 0; w < wid; ++w) before += warp_sum[w];
        uint64_t at = tile_base + before + incl - total;
#pragma unroll
        for (int k = 0; k < kTokPerThread; ++k) {
            image_or(image, at, bits[k], width[k]);
            at += (uint64_t)width[k];
        }
        __syncthreads();
        if (threadIdx.x == kPackThreads - 1) tile_base = at;
        __syncthreads();
    }
    // end-of-block, then the empty stored block that re-aligns the stream: 000, pad, 00 00 FF FF
    if (threadIdx.x == 0) {
        uint64_t at = tile_base;
        const uint32_t eob = lit[256];
        image_or(image, at, eob & 0xFFFFu, (int)(eob >> 16));
        at += eob >> 16;
        at += 3;                       // BFINAL = 0, BTYPE = 00
        at = (at + 7) & ~7ull;         // pad
        image_or(image, at + 16, 0xFFFFu, 16);  // LEN = 0000, NLEN = FFFF
    }
    __syncthreads();

    // image -> stream: bytes up to the first 16-byte boundary, aligned 128-bit stores, byte tail
    {
        const uint8_t *img8 = reinterpret_cast<const uint8_t *>(image);
        uint32_t headn = (16 - lead) & 15;
        if (headn > out_bytes) headn = out_bytes;
        if (threadIdx.x < headn) dest[threadIdx.x] = img8[lead + threadIdx.x];
        const uint32_t nvec = (out_bytes - headn) >> 4;
        const uint4 *sv = reinterpret_cast<const uint4 *>(img8 + lead + headn);
        uint4 *dv = reinterpret_cast<uint4 *>(dest + headn);
        for (uint32_t i = threadIdx.x; i < nvec; i += kPackThreads) dv[i] = sv[i];
        for (uint32_t i = headn + (nvec << 4) + threadIdx.x; i < out_bytes; i += kPackThreads) dest[i] = img8[lead + i];
    }
}

// level <= 0: 65535-byte stored blocks exactly like DeflateStream.compressZlibStored (DS:331-373); the last
// block carries BFINAL when the stream is framed.
__global__ void __launch_bounds__(256) stored_kernel(const uint8_t *__restrict__ in, uint64_t n, uint8_t *__restrict__ out,
                                                     uint32_t num_blocks, int final_flag) {
    for (uint32_t blk = blockIdx.x; blk < num_blocks; blk += gridDim.x) {
        const uint64_t raw0 = (uint64_t)blk * 65535u;
        const uint32_t raw = (uint32_t)tmin<uint64_t>(65535u, n - raw0);
        uint8_t *d = out + (uint64_t)blk * 65540u;
        if (threadIdx.x == 0) {
            d[0] = (uint8_t)((blk == num_blocks - 1) ? final_flag : 0);
            d[1] = (uint8_t)(raw & 0xFF);
            d[2] = (uint8_t)(raw >> 8);
            d[3] = (uint8_t)(~raw & 0xFF);
            d[4] = (uint8_t)((~raw >> 8) & 0xFF);
        }
        const uint8_t *src = in + raw0;
        for (uint32_t i = threadIdx.x; i < raw; i += 256) d[5 + i] = src[i];
    }
}

// zlib header / final empty fixed block / Adler-32 trailer / result record
__global__ void frame_kernel(uint8_t *out, uint64_t cap, int level, int write_header, int write_final_block,
                             const uint64_t *cursor, int64_t body_override, const uint32_t *adler, uint64_t n,
                             uint64_t *result) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint64_t body = body_override >= 0 ? (uint64_t)body_override : cursor[0];
    uint64_t total = body;
    int status = ZK_OK;
    uint8_t *p = out;
    if (write_header) {
        total += 2 + 4 + (write_final_block ? 2 : 0);
        if (cursor[1] != 0 || total > cap) {
            status = ZK_BUF_ERROR;
        } else {
            const int flevel = level >= 9 ? 3 : level >= 6 ? 2 : level >= 2 ? 1 : 0;  // DS:27-33
            const int cmf = 0x78, flg0 = flevel << 6;
            p[0] = (uint8_t)cmf;
            p[1] = (uint8_t)(flg0 | ((31 - (((cmf << 8) | flg0) % 31)) % 31));  // DS:45-47
            uint8_t *t = p + 2 + body;
            if (write_final_block) {
                *t++ = 0x03;  // BFINAL = 1, BTYPE = 01, end-of-block (7 zero bits) spills into ...
                *t++ = 0x00;  // ... this byte; same two bytes the reference emits for empty input
            }
            const uint32_t a = adler[0];
            t[0] = (uint8_t)(a >> 24);  // DS:57-74 big-endian trailer
            t[1] = (uint8_t)(a >> 16);
            t[2] = (uint8_t)(a >> 8);
            t[3] = (uint8_t)a;
        }
    } else if (cursor[1] != 0 || total > cap) {
        status = ZK_BUF_ERROR;
    }
    result[0] = status == ZK_OK ? total : 0;
    result[1] = adler[0];
    result[2] = (uint64_t)(int64_t)status;
    result[3] = n;
}

}  // namespace

size_t pack_image_bytes(uint32_t block_bytes) { return (size_t)block_bytes + kMaxHeaderBytes + 64; }

int launch_pack(zk_ctx *c, const uint8_t *d_in, uint64_t n, uint32_t block_bytes, uint32_t num_blocks,
                const uint32_t *tokens, const uint32_t *block_tok_begin, const uint32_t *block_tok_count,
                const BlockPlan *plans, const uint64_t *offsets, const uint64_t *cursor, uint8_t *out) {
    PackArgs a;
    a.in = d_in;
    a.n = n;
    a.block_bytes = block_bytes;
    a.num_blocks = num_blocks;
    a.tokens = tokens;
    a.block_tok_begin = block_tok_begin;
    a.block_tok_count = block_tok_count;
    a.plans = plans;
    a.offsets = offsets;
    a.cursor = cursor;
    a.out = out;
    const size_t smem = pack_image_bytes(block_bytes);
    a.image_words = (uint32_t)(smem / 4);
    ZK_CUDA(cudaFuncSetAttribute(pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pack_kernel<<<num_blocks, kPackThreads, smem, c->stream>>>(a);
    c->launches += 1;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}

int launch_stored(zk_ctx *c, const uint8_t *d_in, uint64_t n, uint8_t *out, uint32_t num_blocks, int final_flag) {
    const uint32_t grid = num_blocks < (uint32_t)c->sm_count * 8 ? num_blocks : (uint32_t)c->sm_count * 8;
    stored_kernel<<<grid, 256, 0, c->stream>>>(d_in, n, out, num_blocks, final_flag);
    c->launches += 1;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}

int launch_frame(zk_ctx *c, uint8_t *out, uint64_t cap, int level, int write_header, int write_final_block,
                 const uint64_t *cursor, int64_t body_override, const uint32_t *adler, uint64_t n, uint64_t *result) {
    frame_kernel<<<1, 32, 0, c->stream>>>(out, cap, level, write_header, write_final_block, cursor, body_override, adler,
                                          n, result);
    c->launches += 1;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}

}  // namespace zk
