// zk_adler.cu -- K4: Adler-32 as a parallel reduction with the zlib combine step.
//
// Replaces Adler32Utils.adler32 (bitwise/checksum/Adler32Utils.kt:31-57): s1 = seed_lo + sum(b),
// s2 = seed_hi + sum of the running s1, both mod 65521.  For a span X of length L started from
// (0,0):  A_X = sum b_i,  B_X = sum (L - i) b_i.  Spans concatenate as
//     A = A_X + A_Y,   B = B_X + L_Y * A_X + B_Y        (mod 65521)
// which is zlib's adler32_combine identity; the seed is a leading span of length 0.
//
// Roofline: pure HBM read, N algorithmic bytes.  One pass: every thread streams 16-byte vectors
// (ld.global.nc, L1 no-allocate, 4 in flight) and keeps two accumulators; dp4a does the byte sums.
#include "zk_common.cuh"

namespace zk {

namespace {

constexpr uint32_t kBase = 65521u;
constexpr int kThreads = 256;
constexpr int kUnroll = 4;

struct AdlerPartial {
    uint32_t a, b;
    uint64_t len;
};

__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

// bytes of one 16-byte vector: plain sum and sum of k*b_k (k = 0..15)
__device__ __forceinline__ void vec_sums(const uint4 &v, uint32_t &plain, uint32_t &weighted) {
    uint32_t s0 = __dp4a(v.x, 0x01010101u, 0u), s1 = __dp4a(v.y, 0x01010101u, 0u);
    uint32_t s2 = __dp4a(v.z, 0x01010101u, 0u), s3 = __dp4a(v.w, 0x01010101u, 0u);
    uint32_t w = __dp4a(v.x, 0x03020100u, 0u);
    w = __dp4a(v.y, 0x03020100u, w);
    w = __dp4a(v.z, 0x03020100u, w);
    w = __dp4a(v.w, 0x03020100u, w);
    plain = s0 + s1 + s2 + s3;
    weighted = w + 4u * s1 + 8u * s2 + 12u * s3;
}

// Partial 0 = unaligned head bytes, partials 1..grid = 16-byte-aligned spans, partial grid+1 = tail.
__global__ void __launch_bounds__(kThreads) adler_spans_kernel(const uint8_t *__restrict__ buf, uint64_t n,
                                                              AdlerPartial *__restrict__ partials) {
    const uint64_t head = tmin<uint64_t>(n, (16u - (uint32_t)(reinterpret_cast<uintptr_t>(buf) & 15u)) & 15u);
    const uint64_t nvec = (n - head) >> 4;
    const uint64_t tail = n - head - (nvec << 4);
    const uint4 *vec = reinterpret_cast<const uint4 *>(buf + head);

    // scalar edges: done by one thread of the first / last CTA
    if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
        if (blockIdx.x == 0) {
            uint32_t a = 0, b = 0;
            for (uint64_t i = 0; i < head; ++i) {
                a += buf[i];
                b += a;
            }
            partials[0] = {a, b, head};  // head <= 15: no reduction needed
        }
        if (blockIdx.x == gridDim.x - 1) {
            uint32_t a = 0, b = 0;
            const uint8_t *t = buf + head + (nvec << 4);
            for (uint64_t i = 0; i < tail; ++i) {
                a += t[i];
                b += a;
            }
            partials[gridDim.x + 1] = {a, b, tail};
        }
    }

    // this CTA's span of vectors [v0, v1)
    const uint64_t per = (nvec + gridDim.x - 1) / gridDim.x;
    const uint64_t v0 = min(nvec, per * blockIdx.x), v1 = min(nvec, v0 + per);
    const uint64_t span_bytes = (v1 - v0) << 4;

    uint32_t s1 = 0;  // sum of bytes seen by this thread (< 2^32: 4080 per vector)
    uint64_t s2 = 0;  // sum of (span_bytes - offset) * byte
    uint64_t v = v0 + threadIdx.x;
    for (; v + (kUnroll - 1) * kThreads < v1; v += kUnroll * kThreads) {
        uint4 x[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(vec + v + u * kThreads);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            uint32_t plain, weighted;
            vec_sums(x[u], plain, weighted);
            const uint64_t w = span_bytes - ((v + u * kThreads - v0) << 4);  // weight of byte 0 of the vector
            s1 += plain;
            s2 += w * plain - weighted;
        }
    }
    for (; v < v1; v += kThreads) {
        uint32_t plain, weighted;
        vec_sums(ld_stream(vec + v), plain, weighted);
        const uint64_t w = span_bytes - ((v - v0) << 4);
        s1 += plain;
        s2 += w * plain - weighted;
    }

    // CTA reduction (values stay far below 2^64: s2 <= span_bytes * 255 * span_bytes/2)
    uint64_t r1 = s1, r2 = s2 % kBase;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        r1 += __shfl_down_sync(0xFFFFFFFFu, r1, off);
        r2 += __shfl_down_sync(0xFFFFFFFFu, r2, off);
    }
    __shared__ uint64_t w1[kThreads / 32], w2[kThreads / 32];
    if ((threadIdx.x & 31) == 0) {
        w1[threadIdx.x >> 5] = r1;
        w2[threadIdx.x >> 5] = r2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t a = 0, b = 0;
        for (int i = 0; i < kThreads / 32; ++i) {
            a += w1[i];
            b += w2[i];
        }
        partials[1 + blockIdx.x] = {(uint32_t)(a % kBase), (uint32_t)(b % kBase), span_bytes};
    }
}

// Fold the partials in order, starting from the seed.  One warp; the list is a few hundred long.
// One warp: lane l folds a contiguous run of partials (independent loads), the lanes are combined pairwise in order with
// shuffles (the combine is associative), lane 0 puts the seed in front.  The serial loop this replaces took as long as
// the span kernel itself at 1 GiB.
__global__ void adler_fold_kernel(const AdlerPartial *__restrict__ partials, int count, uint32_t seed,
                                  uint32_t *__restrict__ out) {
    if (blockIdx.x != 0) return;
    const int lane = threadIdx.x & 31;
    const int run = (count + 31) / 32;
    uint64_t a = 0, b = 0, len = 0;
    for (int i = lane * run; i < min(count, (lane + 1) * run); ++i) {
        const AdlerPartial p = partials[i];
        b = (b + (p.len % kBase) * a + p.b) % kBase;
        a = (a + p.a) % kBase;
        len += p.len;
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint64_t a2 = __shfl_down_sync(0xFFFFFFFFu, a, off), b2 = __shfl_down_sync(0xFFFFFFFFu, b, off);
        const uint64_t l2 = __shfl_down_sync(0xFFFFFFFFu, len, off);
        if (lane + off < 32) {
            b = (b + (l2 % kBase) * a + b2) % kBase;
            a = (a + a2) % kBase;
            len += l2;
        }
    }
    if (lane == 0) {
        const uint64_t sa = seed & 0xFFFFu, sb = (seed >> 16) & 0xFFFFu;
        const uint64_t rb = (sb + (len % kBase) * sa + b) % kBase, ra = (sa + a) % kBase;
        out[0] = (uint32_t)((rb << 16) | ra);
    }
}

int adler_grid(const zk_ctx *c, uint64_t n) {
    // enough CTAs to fill every SM with 8 resident CTAs, never more than one per 64 KiB of input
    uint64_t want = (uint64_t)c->sm_count * 8;
    uint64_t by_size = (n >> 16) + 1;
    return (int)(want < by_size ? want : by_size);
}

}  // namespace

// (sized for the largest grid adler_grid hands out on any device this can run on: 8 CTAs per SM, up to 256 SMs)
size_t adler_partials_bytes(uint64_t n) { return sizeof(AdlerPartial) * (size_t)(256 * 8 + 2); }

int launch_adler32(zk_ctx *c, uint32_t seed, const uint8_t *d_buf, uint64_t n, void *d_partials, uint32_t *d_out) {
    const int grid = adler_grid(c, n);
    AdlerPartial *p = static_cast<AdlerPartial *>(d_partials);
    adler_spans_kernel<<<grid, kThreads, 0, c->stream>>>(d_buf, n, p);
    adler_fold_kernel<<<1, 32, 0, c->stream>>>(p, grid + 2, seed, d_out);
    c->launches += 2;
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}

// zlib's adler32_combine on three scalars (host; used to fold per-shard checksums).
uint32_t adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2) {
    const uint64_t rem = len2 % kBase;
    uint64_t s1 = a1 & 0xFFFFu, s2 = (a1 >> 16) & 0xFFFFu;
    // a2 was computed from seed 1: remove that seed's contribution (1 to s1, len2 to s2)
    uint64_t t1 = (a2 & 0xFFFFu) + kBase - 1;
    uint64_t t2 = ((a2 >> 16) & 0xFFFFu) + kBase - rem;
    s2 = (s2 + rem * s1 + t2) % kBase;
    s1 = (s1 + t1) % kBase;
    return (uint32_t)((s2 << 16) | s1);
}

}  // namespace zk
