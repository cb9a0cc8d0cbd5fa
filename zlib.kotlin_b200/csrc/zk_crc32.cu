// zk_crc32.cu -- CRC-32 (IEEE 802.3, reflected, polynomial 0xEDB88320) as a parallel reduction: the checksum of the
// gzip container (RFC 1952), SURVEY.md 8(f) row 4.  The reference has no CRC-32 (only a logging stub,
// common/ZlibLogger.kt:115-118); libz's crc32 / crc32_combine are the oracle.
//
// A CRC is linear over GF(2): for raw register states started from zero, raw(A || B) = raw(A) * x^(8|B|) + raw(B) mod P,
// and leading zero bytes leave a zero register at zero.  The buffer is cut into equal pieces that are aligned to its
// END (the first piece may be short: that is the same as zero padding in front of it), every thread runs the
// slice-by-4 table recurrence over its piece, and pieces are folded pairwise with the constants x^(8 * piece * 2^k)
// (computed on the host by square-and-multiply): five shuffle levels inside a warp, three through shared memory, then
// one warp folds the per-CTA values.  The pre/post conditioning of CRC-32 is applied once at the end:
// crc(msg, seed) = ~(raw(msg) + (~seed) * x^(8|msg|)).
//
// Roofline: pure HBM read, N algorithmic bytes.
#include "zk_common.cuh"

namespace zk {

namespace {

constexpr uint32_t kPoly = 0xEDB88320u;
constexpr int kCrcThreads = 256;

struct CrcConsts {
    uint32_t piece[8];  // x^(8 * P * 2^k) for k = 0..7: folds inside a CTA
    uint32_t run[6];    // [0]: x^(8 * 256 P) (one CTA); [1 + k]: x^(8 * 256 P * R * 2^k): folds of the lanes' runs of R CTAs
    uint32_t total;     // x^(8 n): carries the seed over the whole message
};

// a * b mod P, both in the reflected representation (bit 31 = x^0), zlib's multmodp without the early exit
__host__ __device__ __forceinline__ uint32_t mulmod(uint32_t a, uint32_t b) {
    uint32_t p = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        p ^= (a & (0x80000000u >> i)) ? b : 0u;
        b = (b >> 1) ^ ((b & 1u) ? kPoly : 0u);
    }
    return p;
}

__global__ void __launch_bounds__(kCrcThreads) crc_pieces_kernel(const uint8_t *__restrict__ buf, uint64_t main_len,
                                                                 uint64_t piece, CrcConsts k, uint32_t *__restrict__ cta_raw) {
    __shared__ uint32_t tab[4][256];
    __shared__ uint32_t warp_val[kCrcThreads / 32];
    {
        uint32_t c = threadIdx.x;
        for (int b = 0; b < 8; ++b) c = (c >> 1) ^ ((c & 1u) ? kPoly : 0u);
        tab[0][threadIdx.x] = c;
    }
    __syncthreads();
    {
        uint32_t c = tab[0][threadIdx.x];
        for (int t = 1; t < 4; ++t) {
            c = (c >> 8) ^ tab[0][c & 0xFFu];
            tab[t][threadIdx.x] = c;
        }
    }
    __syncthreads();
    // pieces are numbered from the front but aligned to the end of [0, main_len)
    const uint64_t total_pieces = (uint64_t)gridDim.x * kCrcThreads;
    const uint64_t t = (uint64_t)blockIdx.x * kCrcThreads + threadIdx.x;
    const uint64_t back = (total_pieces - t) * piece;  // distance of the piece's start from main_len
    uint64_t lo = back <= main_len ? main_len - back : 0;
    const uint64_t hi = back - piece <= main_len ? main_len - (back - piece) : 0;
    uint32_t c = 0;
    // unaligned head (only the piece that contains offset 0 can have one: all other boundaries are 16-byte aligned)
    while (lo < hi && ((reinterpret_cast<uintptr_t>(buf + lo)) & 15u) != 0) {
        c = tab[0][(c ^ buf[lo]) & 0xFFu] ^ (c >> 8);
        ++lo;
    }
    const uint4 *v = reinterpret_cast<const uint4 *>(buf + lo);
    const uint64_t nvec = (hi - lo) >> 4;
    for (uint64_t i = 0; i < nvec; ++i) {
        const uint4 w = __ldg(v + i);
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            c ^= ws[j];
            c = tab[3][c & 0xFFu] ^ tab[2][(c >> 8) & 0xFFu] ^ tab[1][(c >> 16) & 0xFFu] ^ tab[0][c >> 24];
        }
    }
    for (uint64_t i = lo + (nvec << 4); i < hi; ++i) c = tab[0][(c ^ buf[i]) & 0xFFu] ^ (c >> 8);
    // fold: the lower lane holds the earlier piece
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int lv = 0; lv < 5; ++lv) {
        const uint32_t later = __shfl_down_sync(0xFFFFFFFFu, c, 1 << lv);
        c = mulmod(c, k.piece[lv]) ^ later;
    }
    if (lane == 0) warp_val[warp] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t a = 0;
        for (int w = 0; w < kCrcThreads / 32; ++w) a = mulmod(a, k.piece[5]) ^ warp_val[w];
        cta_raw[blockIdx.x] = a;
    }
}

// one warp: folds the per-CTA values, appends the unaligned tail, applies the seed and the final inversion
__global__ void crc_fold_kernel(const uint32_t *__restrict__ cta_raw, uint32_t num_cta, uint32_t run, CrcConsts k,
                                const uint8_t *__restrict__ tail, uint32_t tail_len, uint32_t seed, uint32_t *out) {
    const int lane = threadIdx.x;
    uint32_t c = 0;
    // lane l folds CTAs [G - (32 - l) * run, G - (31 - l) * run): aligned to the end again
    const int64_t first = (int64_t)num_cta - (int64_t)(32 - lane) * run;
    for (int64_t g = first; g < first + (int64_t)run; ++g)
        if (g >= 0) c = mulmod(c, k.run[0]) ^ cta_raw[g];
#pragma unroll
    for (int lv = 0; lv < 5; ++lv) {
        const uint32_t later = __shfl_down_sync(0xFFFFFFFFu, c, 1 << lv);
        c = mulmod(c, k.run[1 + lv]) ^ later;
    }
    if (lane == 0) {
        for (uint32_t i = 0; i < tail_len; ++i) {
            c ^= tail[i];
            for (int b = 0; b < 8; ++b) c = (c >> 1) ^ ((c & 1u) ? kPoly : 0u);
        }
        out[0] = ~(c ^ mulmod(~seed, k.total));
    }
}

uint32_t xpow8(uint64_t nbytes) {  // x^(8 * nbytes) mod P
    uint32_t r = 0x80000000u, b = 0x40000000u;  // x^0, x^1
    // exponent 8 * nbytes as up to 67 bits: square three times first, then the bits of nbytes
    for (int i = 0; i < 3; ++i) b = mulmod(b, b);  // b = x^8
    for (uint64_t e = nbytes; e; e >>= 1) {
        if (e & 1) r = mulmod(r, b);
        b = mulmod(b, b);
    }
    return r;
}

}  // namespace

uint32_t crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) { return mulmod(xpow8(len2), crc1) ^ crc2; }

size_t crc_partials_bytes(uint64_t n) {
    (void)n;
    return 4096 * sizeof(uint32_t);
}

int launch_crc32(zk_ctx *c, uint32_t seed, const uint8_t *d_buf, uint64_t n, void *d_partials, uint32_t *d_out) {
    // main part ends on a 16-byte aligned address, the rest (< 16 bytes) is folded in by the last kernel
    const uint64_t end_mis = (reinterpret_cast<uintptr_t>(d_buf) + n) & 15u;
    const uint64_t tail_len = tmin<uint64_t>(end_mis, n);
    const uint64_t main_len = n - tail_len;
    uint32_t grid = (uint32_t)tmin<uint64_t>((uint64_t)c->sm_count * 4, (main_len + (64u << 10) - 1) / (64u << 10));
    if (grid == 0) grid = 1;
    if (grid > 4096) grid = 4096;
    const uint64_t pieces = (uint64_t)grid * kCrcThreads;
    uint64_t piece = (main_len + pieces - 1) / pieces;
    piece = (piece + 15) / 16 * 16;
    if (piece == 0) piece = 16;
    const uint32_t run = (grid + 31) / 32;
    CrcConsts k;
    for (int lv = 0; lv < 8; ++lv) k.piece[lv] = xpow8(piece << lv);
    k.run[0] = xpow8(piece * kCrcThreads);
    for (int lv = 0; lv < 5; ++lv) k.run[1 + lv] = xpow8(piece * kCrcThreads * run << lv);
    k.total = xpow8(n);
    uint32_t *d_cta = static_cast<uint32_t *>(d_partials);
    crc_pieces_kernel<<<grid, kCrcThreads, 0, c->stream>>>(d_buf, main_len, piece, k, d_cta);
    crc_fold_kernel<<<1, 32, 0, c->stream>>>(d_cta, grid, run, k, d_buf + main_len, (uint32_t)tail_len, seed, d_out);
    c->launches += 2;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}

}  // namespace zk
