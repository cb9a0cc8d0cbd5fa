// zk_api.cu -- the C ABI of include/zlibk_gpu.h: contexts, orchestration of K1..K5, host-buffer wrappers
// and the streaming handles.  Host code only; every kernel lives in its own translation unit.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <memory>
#include <vector>

#include "zk_common.cuh"

namespace zk {

// launchers defined next to their kernels
int launch_lz77(zk_ctx *c, const uint8_t *d_in, uint64_t n, const uint8_t *d_hist, uint32_t hist_len,
                uint32_t unit_bytes, uint32_t block_bytes, const LevelParams &lp, uint32_t *tokens,
                uint32_t *block_tok_begin, uint32_t *block_tok_count, unsigned long long *prof);
int launch_hist(zk_ctx *c, const uint32_t *tokens, const uint32_t *block_tok_begin, const uint32_t *block_tok_count,
                uint32_t num_blocks, uint32_t *block_hist);
int launch_huffman(zk_ctx *c, const uint32_t *block_hist, uint64_t n, uint32_t block_bytes, uint32_t num_blocks,
                   int force_type, BlockPlan *plans);
int launch_plan_offsets(zk_ctx *c, const BlockPlan *plans, uint32_t num_blocks, uint64_t *offsets, uint64_t *cursor,
                        uint64_t cap);
int launch_pack(zk_ctx *c, const uint8_t *d_in, uint64_t n, uint32_t block_bytes, uint32_t num_blocks,
                const uint32_t *tokens, const uint32_t *block_tok_begin, const uint32_t *block_tok_count,
                const BlockPlan *plans, const uint64_t *offsets, const uint64_t *cursor, uint8_t *out);
int launch_stored(zk_ctx *c, const uint8_t *d_in, uint64_t n, uint8_t *out, uint32_t num_blocks, int final_flag);
int launch_frame(zk_ctx *c, uint8_t *out, uint64_t cap, int level, int write_header, int write_final_block,
                 const uint64_t *cursor, int64_t body_override, const uint32_t *adler, uint64_t n, uint64_t *result);

static thread_local char g_last_error[256] = "";

void set_last_error(const char *what, cudaError_t e) {
    snprintf(g_last_error, sizeof g_last_error, "%s: %s", what, cudaGetErrorString(e));
}

// ---- optional device timing per kernel class -------------------------------------------------------------
ScopedTimer::ScopedTimer(zk_ctx *ctx, int cls) : c(ctx), idx(-1) {
    if (!c->opt_timing) return;
    if (c->span_count == c->span_cap) {
        const int ncap = c->span_cap ? c->span_cap * 2 : 64;
        zk_ctx::TimedSpan *n = static_cast<zk_ctx::TimedSpan *>(realloc(c->spans, sizeof(zk_ctx::TimedSpan) * ncap));
        if (!n) return;
        for (int i = c->span_cap; i < ncap; ++i) {
            cudaEventCreate(&n[i].a);
            cudaEventCreate(&n[i].b);
        }
        c->spans = n;
        c->span_cap = ncap;
    }
    idx = c->span_count++;
    c->spans[idx].cls = cls;
    cudaEventRecord(c->spans[idx].a, c->stream);
}
ScopedTimer::~ScopedTimer() {
    if (idx >= 0) cudaEventRecord(c->spans[idx].b, c->stream);
}
void collect_timers(zk_ctx *c) {
    for (int i = 0; i < c->span_count; ++i) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->spans[i].a, c->spans[i].b) == cudaSuccess) {
            c->class_ms[c->spans[i].cls] += ms;
            c->class_launches[c->spans[i].cls] += 1;
        }
    }
    c->span_count = 0;
    cudaGetLastError();
}

// ---- arena ---------------------------------------------------------------------------------------------
int Arena::reserve(size_t bytes) {
    if (bytes <= cap) return ZK_OK;
    if (base) cudaFree(base);
    base = nullptr;
    cap = 0;
    const size_t want = bytes + (bytes >> 3);
    cudaError_t e = cudaMalloc(&base, want);
    if (e != cudaSuccess) {
        e = cudaMalloc(&base, bytes);
        if (e != cudaSuccess) {
            set_last_error("cudaMalloc(arena)", e);
            base = nullptr;
            return ZK_MEM_ERROR;
        }
        cap = bytes;
        return ZK_OK;
    }
    cap = want;
    return ZK_OK;
}
void *Arena::take(size_t bytes, size_t align) {
    size_t at = (used + align - 1) / align * align;
    if (at + bytes > cap) return nullptr;
    used = at + bytes;
    return base + at;
}
void Arena::release() {
    if (base) cudaFree(base);
    base = nullptr;
    cap = used = 0;
}

// ---- deflate orchestration ---------------------------------------------------------------------------------
enum UnitMode { kUnitSingle, kUnitTailFilled, kUnitLastOfMany };
static uint32_t pick_unit_bytes(const zk_ctx *c, uint64_t n, uint32_t block_bytes, UnitMode mode = kUnitSingle) {
    if (c->opt_chunk_bytes > 0) {
        uint64_t u = (uint64_t)c->opt_chunk_bytes / block_bytes * block_bytes;
        return (uint32_t)tmax<uint64_t>(u, block_bytes);
    }
    // Units fitted to the machine: n cut into waves x SMs equal units (a multiple of the block span).  A unit pays for its
    // 32 KiB of history and for filling its software pipeline once, and a whole number of waves leaves no SM idle at the end
    // of the kernel -- measured at 1 GiB (tools/unit_sweep.py): 512 KiB units 51.0 GB/s, 1 184 KiB (three waves per 512 MiB slab)
    // 52.2, 1 792 KiB (two) 52.4, 3 552 KiB (one) 52.8, but 2 048 KiB (1.7 waves) 49.3.
    const uint64_t sms = (uint64_t)c->sm_count;
    auto fitted = [&](uint64_t waves) {
        uint64_t v = (n + sms * waves - 1) / (sms * waves);
        v = (v + block_bytes - 1) / block_bytes * block_bytes;
        return tmin<uint64_t>(tmax<uint64_t>(v, block_bytes), 8u << 20);
    };
    uint64_t u = 512u << 10;
    if (mode == kUnitTailFilled) {
        // the next slab's LZ kernel takes over an SM as soon as its unit is done: one unit per SM, no tail to keep short
        if (fitted(1) >= (1u << 20)) return (uint32_t)fitted(1);
        // (small slabs: large units, as long as two slabs in flight cover every SM)
        while (u > block_bytes && n / u < (uint64_t)c->sm_count / 2) u >>= 1;
        return (uint32_t)tmax<uint64_t>(u, block_bytes);
    }
    // nothing fills the tail of a single or last slab: two waves, so that units of different cost (long repeats take twice as
    // long as text) even out -- from the size on where such a unit is worth its fixed cost
    if (fitted(2) >= (1u << 20)) return (uint32_t)fitted(2);
    if (mode == kUnitLastOfMany) {
        // short last slab of a pipelined job: fewest waves x (unit + what a unit pays for its 32 KiB of history and for
        // filling its software pipeline, ~52 KiB worth of steps)
        uint64_t best = u, best_cost = ~0ull;
        for (uint64_t v = u; v >= tmax<uint64_t>(block_bytes, 32u << 10); v >>= 1) {
            const uint64_t units = (n + v - 1) / v;
            const uint64_t cost = ((units + c->sm_count - 1) / c->sm_count) * (v + (52u << 10));
            if (cost < best_cost) best = v, best_cost = cost;
        }
        return (uint32_t)tmax<uint64_t>(best, block_bytes);
    }
    // at least ~4 units per SM so the tail wave is short (units differ a lot in cost: a unit of long repeats takes
    // twice as long as a unit of text); history re-priming costs 32 KiB per unit
    while (u > block_bytes && n / u < (uint64_t)c->sm_count * 4) u >>= 1;
    return (uint32_t)tmax<uint64_t>(u, block_bytes);
}

constexpr uint64_t kSlabBytes = 1024ull << 20;  // tokens for one slab: 4 B per input byte (4 GiB of scratch at most)

int run_deflate(zk_ctx *c, const DeflateJob &job) {
    const uint64_t n = job.n;
    const int level = job.level;
    const uint32_t block_bytes = (uint32_t)c->opt_block_bytes;
    uint8_t *body = job.d_out + (job.framed ? 2 : 0);
    const uint64_t frame_bytes = job.framed ? 2 + 2 + 4 : 0;
    const uint64_t body_cap = job.cap > frame_bytes ? job.cap - frame_bytes : 0;

    c->arena.reset();
    c->shard.valid = false;  // the scratch of a pending sharded inflate is gone
    const bool stored_only = level <= 0;
    // A device-resident job of 512 MiB .. 2 GiB runs as two equal slabs: the LZ kernel of the second fills the tail wave
    // of the first, and the histogram / Huffman / pack kernels of the first fill the tail wave of the second (1 GiB:
    // 36.1 -> 37.4 GB/s, profiles/e2e_sweep_r1.md; more, smaller slabs measured slightly worse).
    uint64_t own_slab = kSlabBytes;
    if (c->opt_slab_bytes)
        own_slab = (uint64_t)c->opt_slab_bytes;
    else if (n >= (512ull << 20) && n < 2 * kSlabBytes && c->opt_slab_overlap)
        own_slab = ((n + 1) / 2 + (1u << 20) - 1) >> 20 << 20;
    const uint64_t slab_bytes = job.slab_bytes ? tmin<uint64_t>(job.slab_bytes, kSlabBytes) : own_slab;
    const uint64_t first_bytes = (job.first_slab_bytes && job.first_slab_bytes < slab_bytes) ? job.first_slab_bytes : slab_bytes;
    const uint64_t slab = tmin<uint64_t>(n, slab_bytes);
    const uint32_t slab_blocks = (uint32_t)((slab + block_bytes - 1) / block_bytes) + 1;
    const uint64_t nslabs = n <= first_bytes ? 1 : 1 + (n - first_bytes + slab_bytes - 1) / slab_bytes;
    // two sets of scratch when there is more than one slab: slab k+1 starts while slab k is still running (below)
    const int nsets = (!stored_only && nslabs >= 2 && c->opt_slab_overlap) ? 2 : 1;
    size_t need = 8192 + adler_partials_bytes(n) + (size_t)nslabs * 8 + 256;
    if (!stored_only)
        need += ((size_t)slab * 4 + (size_t)slab_blocks * (8 + kHistStride * 4 + sizeof(BlockPlan) + 8) + 4096) * nsets;
    if (c->arena.reserve(need) != ZK_OK) return ZK_MEM_ERROR;
    uint64_t *d_cursor = static_cast<uint64_t *>(c->arena.take(256));
    uint32_t *d_adler = static_cast<uint32_t *>(c->arena.take(256));
    void *d_partials = c->arena.take(adler_partials_bytes(n));
    ZK_CUDA(cudaMemsetAsync(d_cursor, 0, 16, c->stream));

    int rc;
    if (stored_only) {
        if (job.before_slab)
            for (uint32_t k = 0; (k ? first_bytes + (uint64_t)(k - 1) * slab_bytes : 0) < n; ++k) job.before_slab(job.user, k);
        {
            ScopedTimer t(c, 3);
            rc = launch_adler32(c, 1u, job.d_in, n, d_partials, d_adler);
        }
        if (rc != ZK_OK) return rc;
        // DS:331-373: 65535-byte stored blocks; an input that is a multiple of 65535 (0 included) ends with an
        // empty final block, exactly as the reference's loop does.
        const uint32_t nb = (uint32_t)(n / 65535u) + 1;
        const uint64_t body_len = n + 5ull * nb;
        if (body_len > body_cap) {
            ZK_CUDA(cudaMemsetAsync(d_cursor + 1, 1, 1, c->stream));
        } else {
            rc = launch_stored(c, job.d_in, n, body, nb, job.framed ? 1 : 0);
            if (rc != ZK_OK) return rc;
        }
        return launch_frame(c, job.d_out, job.cap, level, job.framed, 0, d_cursor, (int64_t)body_len, d_adler, n,
                            job.d_result);
    }

    LevelParams lp = level_params(level, job.strategy);
    if (c->opt_lz_chain >= 0) lp.chain = (int)c->opt_lz_chain;
    if (c->opt_lz_max_lazy >= 0) lp.max_lazy = (int)c->opt_lz_max_lazy;
    if (c->opt_lz_nice >= 0) lp.nice = (int)c->opt_lz_nice;
    struct Scratch {
        uint32_t *tokens, *tok_begin, *tok_count, *hist;
        BlockPlan *plans;
        uint64_t *offsets;
    } sets[2];
    for (int i = 0; i < nsets; ++i) {
        sets[i].tokens = static_cast<uint32_t *>(c->arena.take((size_t)slab * 4));
        sets[i].tok_begin = static_cast<uint32_t *>(c->arena.take((size_t)slab_blocks * 4));
        sets[i].tok_count = static_cast<uint32_t *>(c->arena.take((size_t)slab_blocks * 4));
        sets[i].hist = static_cast<uint32_t *>(c->arena.take((size_t)slab_blocks * kHistStride * 4));
        sets[i].plans = static_cast<BlockPlan *>(c->arena.take((size_t)slab_blocks * sizeof(BlockPlan)));
        sets[i].offsets = static_cast<uint64_t *>(c->arena.take((size_t)slab_blocks * 8));
        if (!sets[i].offsets) return ZK_MEM_ERROR;
    }
    uint64_t *d_snap = static_cast<uint64_t *>(c->arena.take((size_t)nslabs * 8));  // the cursor after each slab
    if (!d_snap) return ZK_MEM_ERROR;

    // Slabs alternate between the caller's stream and alt_stream.  Everything of a slab is ordered on its stream; the
    // only cross-slab dependency is the output cursor, which plan_offsets of slab k+1 continues from slab k.  The LZ
    // kernel holds one CTA per SM, so the next slab's CTAs are dispatched exactly as the current slab's last units
    // finish: no SM idles in a slab's tail wave, and the units of a slab that has a successor can be large.
    struct Fork {
        zk_ctx *c;
        cudaStream_t main;
        bool on;
        ~Fork() {
            c->stream = main;
            if (on && cudaEventRecord(c->slab_ev[3], c->alt_stream) == cudaSuccess)
                cudaStreamWaitEvent(main, c->slab_ev[3], 0);  // later work on the caller's stream sees every slab done
        }
    } fork{c, c->stream, false};
    if (nsets == 2) {
        if (!c->alt_stream) ZK_CUDA(cudaStreamCreateWithFlags(&c->alt_stream, cudaStreamNonBlocking));
        for (cudaEvent_t &e : c->slab_ev)
            if (!e) ZK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ZK_CUDA(cudaEventRecord(c->slab_ev[2], c->stream));
        ZK_CUDA(cudaStreamWaitEvent(c->alt_stream, c->slab_ev[2], 0));
        fork.on = true;
    }

    uint32_t slab_index = 0;
    for (uint64_t s0 = 0; s0 < n; s0 += (slab_index ? slab_bytes : first_bytes), ++slab_index) {
        const uint64_t sn = tmin<uint64_t>(slab_index ? slab_bytes : first_bytes, n - s0);
        const Scratch &w = sets[slab_index % nsets];
        c->stream = (slab_index % nsets) ? c->alt_stream : fork.main;
        if (job.before_slab) job.before_slab(job.user, slab_index);
        const uint32_t nb = (uint32_t)((sn + block_bytes - 1) / block_bytes);
        const uint32_t unit = pick_unit_bytes(c, sn, block_bytes, nsets == 1 ? kUnitSingle : s0 + sn < n ? kUnitTailFilled : kUnitLastOfMany);
        const uint8_t *hist = s0 ? job.d_in + s0 - kWindow : job.d_hist;
        const uint32_t hist_len = s0 ? (uint32_t)kWindow : (job.d_hist ? job.hist_len : 0);
        {
            ScopedTimer t(c, 0);
            rc = launch_lz77(c, job.d_in + s0, sn, hist, hist_len, unit, block_bytes, lp, w.tokens, w.tok_begin,
                             w.tok_count, c->lz_prof);
        }
        if (rc != ZK_OK) return rc;
        if (c->lz_dump_tokens && c->lz_dump_blocks && nslabs == 1) {  // tests compare the token stream with the CPU model
            ZK_CUDA(cudaMemcpyAsync(c->lz_dump_tokens, w.tokens, (size_t)sn * 4, cudaMemcpyDeviceToDevice, c->stream));
            ZK_CUDA(cudaMemcpyAsync(c->lz_dump_blocks, w.tok_begin, (size_t)nb * 4, cudaMemcpyDeviceToDevice, c->stream));
            ZK_CUDA(cudaMemcpyAsync(c->lz_dump_blocks + nb, w.tok_count, (size_t)nb * 4, cudaMemcpyDeviceToDevice, c->stream));
        }
        {
            ScopedTimer t(c, 1);
            rc = launch_hist(c, w.tokens, w.tok_begin, w.tok_count, nb, w.hist);
            if (rc == ZK_OK) rc = launch_huffman(c, w.hist, sn, block_bytes, nb, -1, w.plans);
            if (rc == ZK_OK && nsets == 2 && slab_index)
                ZK_CUDA(cudaStreamWaitEvent(c->stream, c->slab_ev[(slab_index - 1) & 1], 0));
            // offsets continue from the previous slab: the cursor lives on the device, no host round trip
            if (rc == ZK_OK) rc = launch_plan_offsets(c, w.plans, nb, w.offsets, d_cursor, body_cap);
            if (rc == ZK_OK) {
                ZK_CUDA(cudaMemcpyAsync(d_snap + slab_index, d_cursor, 8, cudaMemcpyDeviceToDevice, c->stream));
                if (nsets == 2) ZK_CUDA(cudaEventRecord(c->slab_ev[slab_index & 1], c->stream));
            }
        }
        if (rc != ZK_OK) return rc;
        {
            ScopedTimer t(c, 2);
            rc = launch_pack(c, job.d_in + s0, sn, block_bytes, nb, w.tokens, w.tok_begin, w.tok_count, w.plans,
                             w.offsets, d_cursor, body);
        }
        if (rc != ZK_OK) return rc;
        if (job.after_slab) job.after_slab(job.user, slab_index, d_snap + slab_index);
    }
    c->stream = fork.main;
    if (fork.on) {
        fork.on = false;
        ZK_CUDA(cudaEventRecord(c->slab_ev[3], c->alt_stream));
        ZK_CUDA(cudaStreamWaitEvent(c->stream, c->slab_ev[3], 0));
    }
    // the checksum reads the whole input once more (0.2 ms per GiB); it runs last so that a pipelined host copy of the
    // input only has to be complete here
    {
        ScopedTimer t(c, 3);
        rc = launch_adler32(c, 1u, job.d_in, n, d_partials, d_adler);
    }
    if (rc != ZK_OK) return rc;
    return launch_frame(c, job.d_out, job.cap, level, job.framed, job.framed, d_cursor, -1, d_adler, n, job.d_result);
}

}  // namespace zk

// ========================================================================================================
// C ABI
// ========================================================================================================
using namespace zk;

#pragma GCC visibility push(default)
extern "C" {

const char *zk_version(void) { return "1.0.2+b200"; }

int zk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char *zk_last_error(void) { return g_last_error; }

zk_ctx *zk_ctx_create(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) {
        set_last_error("zk_ctx_create: no such CUDA device", e == cudaSuccess ? cudaErrorInvalidDevice : e);
        cudaGetLastError();
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    zk_ctx *c = new (std::nothrow) zk_ctx();
    if (!c) return nullptr;
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return nullptr;
    }
    c->stream = c->own_stream;
    c->pinned_cap = 4096;
    if (cudaHostAlloc(&c->pinned, c->pinned_cap, cudaHostAllocDefault) != cudaSuccess) c->pinned = nullptr;
    return c;
}

static void staging_release(zk_ctx *c);

void zk_ctx_destroy(zk_ctx *c) {
    if (!c) return;
    for (zk_ctx *&h : c->helpers) {
        if (h) zk_ctx_destroy(h);
        h = nullptr;
    }
    for (int i = 0; i < c->num_registered; ++i)
        if (cudaHostUnregister(c->registered[i]) != cudaSuccess) cudaGetLastError();
    c->num_registered = 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    staging_release(c);
    c->arena.release();
    if (c->pinned) cudaFreeHost(c->pinned);
    for (int i = 0; i < c->span_cap; ++i) {
        cudaEventDestroy(c->spans[i].a);
        cudaEventDestroy(c->spans[i].b);
    }
    free(c->spans);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_in) cudaStreamDestroy(c->copy_in);
    if (c->copy_out) cudaStreamDestroy(c->copy_out);
    if (c->alt_stream) cudaStreamDestroy(c->alt_stream);
    for (cudaEvent_t e : c->slab_ev)
        if (e) cudaEventDestroy(e);
    delete c;
}

int zk_ctx_set_stream(zk_ctx *c, void *cuda_stream) {
    if (!c) return ZK_STREAM_ERROR;
    c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
    return ZK_OK;
}

int zk_ctx_synchronize(zk_ctx *c) {
    if (!c) return ZK_STREAM_ERROR;
    ZK_CUDA(cudaStreamSynchronize(c->stream));
    return ZK_OK;
}

int zk_ctx_set_option(zk_ctx *c, const char *name, int64_t value) {
    if (!c || !name) return ZK_STREAM_ERROR;
    if (!strcmp(name, "chunk_bytes")) {
        if (value < 0) return ZK_STREAM_ERROR;
        c->opt_chunk_bytes = value;
        return ZK_OK;
    }
    if (!strcmp(name, "lz_profile_buffer")) {  // device pointer to u64[units][16] cycle counters, 0 = off
        c->lz_prof = reinterpret_cast<unsigned long long *>(static_cast<uintptr_t>(value));
        return ZK_OK;
    }
    if (!strcmp(name, "lz_dump_tokens") || !strcmp(name, "lz_dump_blocks")) {  // device pointers, 0 = off (tests)
        (name[8] == 't' ? c->lz_dump_tokens : c->lz_dump_blocks) = reinterpret_cast<uint32_t *>(static_cast<uintptr_t>(value));
        return ZK_OK;
    }
    if (!strcmp(name, "lz_chain") || !strcmp(name, "lz_max_lazy") || !strcmp(name, "lz_nice")) {  // tuning (tools/lz_profile.py)
        if (value < -1 || value > 258) return ZK_STREAM_ERROR;
        (name[3] == 'c' ? c->opt_lz_chain : name[3] == 'm' ? c->opt_lz_max_lazy : c->opt_lz_nice) = value;
        return ZK_OK;
    }
    if (!strcmp(name, "slab_overlap")) {  // 0: slabs of a multi-slab deflate run strictly one after the other
        c->opt_slab_overlap = value != 0;
        return ZK_OK;
    }
    if (!strcmp(name, "slab_bytes")) {  // slab of a device-resident deflate, 0 = built-in
        if (value < 0 || value > (int64_t)kSlabBytes || (value && (value < (1 << 20) || value % 65536))) return ZK_STREAM_ERROR;
        c->opt_slab_bytes = value;
        return ZK_OK;
    }
    if (!strcmp(name, "host_slab_bytes") || !strcmp(name, "host_first_slab_bytes")) {  // 0 = built-in
        // at most one device slab (run_deflate clamps to that: a larger host slab would index past its per-slab events)
        if (value < 0 || value > (int64_t)kSlabBytes || (value && (value < (1 << 20) || value % 65536))) return ZK_STREAM_ERROR;
        (name[5] == 's' ? c->opt_host_slab : c->opt_host_first_slab) = value;
        return ZK_OK;
    }
    if (!strcmp(name, "auto_register")) {  // 1: page-lock unknown caller buffers on first use and keep them (see the header)
        c->opt_auto_register = value != 0;
        return ZK_OK;
    }
    if (!strcmp(name, "inflate_foreign")) {  // 0: streams without sync markers are decoded by one warp (default 1: entry points
        c->opt_foreign_path = value != 0;    // are searched at dynamic block headers, zk_inflate.cu block_find_kernel)
        return ZK_OK;
    }
    if (!strcmp(name, "inflate_pipe")) {  // 1: large marked streams go through the slab pipeline (default 0: one pass)
        c->opt_inflate_pipe = value != 0;
        return ZK_OK;
    }
    if (!strcmp(name, "inflate_prefix_bytes")) {  // stream handle: a stream without markers is decoded up to its last entry point
        if (value < 0 || (value && value < (256 << 10))) return ZK_STREAM_ERROR;  // whenever this much has gathered (0 = 32 MiB; huge = never)
        c->opt_prefix_bytes = value;
        return ZK_OK;
    }
    if (!strcmp(name, "inflate_slab_bytes")) {  // compressed bytes per slab of the pipelined host-buffer inflate, 0 = built-in
        if (value < 0 || (value && value < (1 << 20))) return ZK_STREAM_ERROR;
        c->opt_inflate_slab = value;
        return ZK_OK;
    }
    if (!strcmp(name, "timing")) {
        c->opt_timing = value != 0;
        return ZK_OK;
    }
    if (!strcmp(name, "block_bytes")) {
        // power of two, a multiple of the LZ group, at most one stored block (<= 65535 bytes)
        if (value < kGroup || value > 32768 || (value & (value - 1))) return ZK_STREAM_ERROR;
        c->opt_block_bytes = value;
        return ZK_OK;
    }
    return ZK_STREAM_ERROR;
}

int64_t zk_ctx_get_counter(zk_ctx *c, const char *name) {
    if (!c || !name) return -1;
    if (!strcmp(name, "kernel_launches")) return c->launches;
    if (!strcmp(name, "h2d_bytes")) return c->h2d_bytes;
    if (!strcmp(name, "d2h_bytes")) return c->d2h_bytes;
    if (!strcmp(name, "inflate_foreign_chunks")) return c->foreign_chunks;
    if (!strcmp(name, "inflate_foreign_merges")) return c->foreign_merges;
    if (!strcmp(name, "inflate_quick_incomplete")) return c->quick_incomplete;
    if (!strcmp(name, "inflate_prefix_decodes")) return c->prefix_decodes;
    if (!strcmp(name, "inflate_pipe_slabs")) return c->pipe_slabs;
    if (!strcmp(name, "inflate_pipe_fallbacks")) return c->pipe_fallbacks;
    // device time per kernel class in microseconds / launches timed; call after zk_ctx_synchronize
    static const char *cls[] = {"lz77", "huffman", "pack", "adler", "inflate_decode", "inflate_other"};
    for (int i = 0; i < zk_ctx::kTimedClasses; ++i) {
        char key[48];
        snprintf(key, sizeof key, "%s_us", cls[i]);
        if (!strcmp(name, key)) {
            collect_timers(c);
            return (int64_t)(c->class_ms[i] * 1000.0);
        }
        snprintf(key, sizeof key, "%s_spans", cls[i]);
        if (!strcmp(name, key)) {
            collect_timers(c);
            return c->class_launches[i];
        }
    }
    if (!strcmp(name, "reset_timing")) {
        collect_timers(c);
        for (int i = 0; i < zk_ctx::kTimedClasses; ++i) {
            c->class_ms[i] = 0;
            c->class_launches[i] = 0;
        }
        return 0;
    }
    return -1;
}

uint32_t zk_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2) { return adler32_combine(a1, a2, len2); }

// 10 bytes per stored block (header + sync marker) of the smallest block span the "block_bytes" option allows (4 KiB),
// + framing: holds for every block_bytes, on incompressible data too
uint64_t zk_compress_bound(uint64_t n) { return n + 10 * (n / 4096 + 1) + 64; }

void zk_zlib_header(int level, uint8_t out2[2]) {
    const int flevel = level >= 9 ? 3 : level >= 6 ? 2 : level >= 2 ? 1 : 0;  // DS:27-33
    const int cmf = 0x78, flg = flevel << 6;
    out2[0] = (uint8_t)cmf;
    out2[1] = (uint8_t)(flg | ((31 - (((cmf << 8) | flg) % 31)) % 31));
}

void zk_zlib_tail(uint32_t adler, uint8_t out6[6]) {
    out6[0] = 0x03;
    out6[1] = 0x00;
    out6[2] = (uint8_t)(adler >> 24);
    out6[3] = (uint8_t)(adler >> 16);
    out6[4] = (uint8_t)(adler >> 8);
    out6[5] = (uint8_t)adler;
}

// ---- device-buffer entry points ------------------------------------------------------------------------------
int zk_adler32_device(zk_ctx *c, uint32_t adler, const uint8_t *d_buf, uint64_t len, uint32_t *d_out) {
    if (!c || !d_out || (!d_buf && len)) return ZK_STREAM_ERROR;
    ZK_CUDA(cudaSetDevice(c->device));
    c->arena.reset();
    c->shard.valid = false;  // the scratch of a pending sharded inflate is gone
    if (c->arena.reserve(adler_partials_bytes(len) + 1024) != ZK_OK) return ZK_MEM_ERROR;
    void *partials = c->arena.take(adler_partials_bytes(len));
    return launch_adler32(c, adler, d_buf, len, partials, d_out);
}

int zk_deflate_zlib_device(zk_ctx *c, const uint8_t *d_in, uint64_t n, int level, int strategy, uint8_t *d_out,
                           uint64_t cap, uint64_t *d_result) {
    if (!c || !d_out || !d_result || (!d_in && n)) return ZK_STREAM_ERROR;
    if (level > 9) return ZK_STREAM_ERROR;  // DS:323-325
    if (strategy < 0 || strategy > ZK_HUFFMAN_ONLY) return ZK_STREAM_ERROR;
    ZK_CUDA(cudaSetDevice(c->device));
    DeflateJob job{d_in, n, nullptr, 0, level, strategy, d_out, cap, true, d_result};
    return run_deflate(c, job);
}

int zk_deflate_shard_device(zk_ctx *c, const uint8_t *d_hist, uint32_t hist_len, const uint8_t *d_in, uint64_t n,
                            int level, int strategy, uint8_t *d_body, uint64_t cap, uint64_t *d_meta) {
    if (!c || !d_body || !d_meta || (!d_in && n) || hist_len > (uint32_t)kWindow) return ZK_STREAM_ERROR;
    if (level > 9) return ZK_STREAM_ERROR;
    if (strategy < 0 || strategy > ZK_HUFFMAN_ONLY) return ZK_STREAM_ERROR;
    ZK_CUDA(cudaSetDevice(c->device));
    DeflateJob job{d_in, n, hist_len ? d_hist : nullptr, hist_len, level, strategy, d_body, cap, false, d_meta};
    return run_deflate(c, job);
}

int zk_inflate_zlib_device(zk_ctx *c, const uint8_t *d_in, uint64_t m, uint8_t *d_out, uint64_t cap,
                           uint64_t *d_result) {
    if (!c || !d_result || (!d_in && m) || (!d_out && cap)) return ZK_STREAM_ERROR;
    ZK_CUDA(cudaSetDevice(c->device));
    InflateJob job{d_in, m, d_out, cap, d_result};
    return run_inflate(c, job);
}

// ---- one stream inflated across GPUs ("Partitioning (inflate)") -------------------------------------
// Every rank holds the body of a chunk range (what zk_deflate_shard_device made).  Phase 1 decodes it with references
// into the preceding ranks' output kept symbolic and hands out the shard's total tail map (u16[32768]: slot = a byte of
// the shard's last 32 KiB, or 0x8000 | index into the window in front of the shard).  The ranks exchange these maps with
// ONE all_gather (64 KiB each); zk_window_from_maps_device composes the maps of the ranks before this one into the 32 KiB
// in front of its shard; phase 2 resolves the shard with it and writes the bytes.  d_result = {out_len, consumed, status,
// Adler-32 of the shard's output} -- fold the checksums with zk_adler32_combine and compare with the stream's trailer.
int zk_inflate_shard_prepare_device(zk_ctx *c, const uint8_t *d_body, uint64_t m, uint32_t hist_len, uint64_t cap,
                                    uint16_t *d_tail_map, uint64_t *d_result) {
    if (!c || !d_tail_map || !d_result || (!d_body && m) || hist_len > (uint32_t)kWindow) return ZK_STREAM_ERROR;
    ZK_CUDA(cudaSetDevice(c->device));
    InflateJob job{d_body, m, nullptr, cap, d_result};
    job.container = 3;
    job.dict_len = hist_len;
    job.d_tail_map = d_tail_map;
    job.quick_incomplete = c->stream_quick;
    c->shard.valid = false;
    return run_inflate(c, job);
}

int zk_window_from_maps_device(zk_ctx *c, const uint16_t *d_maps, uint64_t stride_bytes, uint32_t count, uint8_t *d_window) {
    if (!c || !d_window || (!d_maps && count) || (stride_bytes & 1)) return ZK_STREAM_ERROR;
    ZK_CUDA(cudaSetDevice(c->device));
    return launch_window_from_maps(c, d_maps, stride_bytes / 2, count, d_window);
}

int zk_inflate_shard_finish_device(zk_ctx *c, const uint8_t *d_entry_window, uint8_t *d_out, uint64_t *d_result) {
    if (!c || !d_entry_window || !d_result) return ZK_STREAM_ERROR;
    ZK_CUDA(cudaSetDevice(c->device));
    return run_inflate_shard_finish(c, d_entry_window, d_out, d_result);
}

// ---- host-buffer entry points ----------------------------------------------------------------------------------
// Device staging for host calls: one grow-only buffer per context, separate from the kernel arena.
struct Staging {
    uint8_t *base = nullptr;
    size_t cap = 0;
};
static thread_local std::vector<std::pair<zk_ctx *, Staging>> g_staging;

static Staging *staging_for(zk_ctx *c, size_t bytes) {
    Staging *s = nullptr;
    for (auto &p : g_staging)
        if (p.first == c) s = &p.second;
    if (!s) {
        g_staging.push_back({c, Staging()});
        s = &g_staging.back().second;
    }
    if (s->cap < bytes) {
        if (s->base) cudaFree(s->base);
        s->base = nullptr;
        s->cap = 0;
        if (cudaMalloc(&s->base, bytes + (bytes >> 4) + 4096) != cudaSuccess) {
            cudaGetLastError();
            if (cudaMalloc(&s->base, bytes) != cudaSuccess) {
                set_last_error("cudaMalloc(staging)", cudaErrorMemoryAllocation);
                return nullptr;
            }
            s->cap = bytes;
        } else {
            s->cap = bytes + (bytes >> 4) + 4096;
        }
    }
    return s;
}

// the calling thread's staging buffer of a context that is being destroyed
static void staging_release(zk_ctx *c) {
    for (size_t i = 0; i < g_staging.size(); ++i)
        if (g_staging[i].first == c) {
            if (g_staging[i].second.base) cudaFree(g_staging[i].second.base);
            g_staging.erase(g_staging.begin() + (long)i);
            return;
        }
}

// "auto_register": page-lock a caller buffer CUDA does not know yet, once, and remember it (the caller keeps it alive and
// reuses it -- the contract of the option; include/zlibk_gpu.h)
static void auto_register(zk_ctx *c, const void *p, uint64_t bytes) {
    if (!c->opt_auto_register || !p || bytes < (1u << 20)) return;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    if (at.type != cudaMemoryTypeUnregistered || c->num_registered >= 16) return;
    if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    c->registered[c->num_registered++] = const_cast<void *>(p);
}

void *zk_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (!bytes || cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void zk_host_free(void *p) {
    if (p) cudaFreeHost(p);
}
int zk_host_register(void *p, uint64_t bytes) {
    if (!p || !bytes) return ZK_STREAM_ERROR;
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return ZK_OK;
    }
    if (e != cudaSuccess) {
        set_last_error("cudaHostRegister", e);
        cudaGetLastError();
        return ZK_MEM_ERROR;
    }
    return ZK_OK;
}
int zk_host_unregister(void *p) {
    if (!p) return ZK_STREAM_ERROR;
    if (cudaHostUnregister(p) != cudaSuccess) {
        cudaGetLastError();
        return ZK_STREAM_ERROR;
    }
    return ZK_OK;
}

uint32_t zk_adler32(zk_ctx *c, uint32_t adler, const uint8_t *buf, uint64_t len) {
    if (buf == nullptr) return 1u;  // AU:37
    if (!c || len == 0) return len == 0 ? adler : 0u;
    if (cudaSetDevice(c->device) != cudaSuccess) return 0u;
    Staging *s = staging_for(c, (size_t)len + 256);
    if (!s) return 0u;
    uint32_t *d_out = reinterpret_cast<uint32_t *>(s->base);
    uint8_t *d_buf = s->base + 256;
    if (cudaMemcpyAsync(d_buf, buf, len, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return 0u;
    c->h2d_bytes += (int64_t)len;
    if (zk_adler32_device(c, adler, d_buf, len, d_out) != ZK_OK) return 0u;
    uint32_t out = 0;
    if (cudaMemcpyAsync(&out, d_out, 4, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return 0u;
    cudaStreamSynchronize(c->stream);
    c->d2h_bytes += 4;
    return out;
}

// Pipelined host call: the input goes up in slabs on its own stream, slab k is compressed as soon as it has arrived,
// and the bytes slab k produced come down on a third stream while slab k+1 is being compressed.
namespace {
// Slab sizes, measured at 1 GiB with the slabs overlapping on two streams (tools/e2e_sweep.py, profiles/e2e_sweep_r1.md):
// small slabs cost nothing any more (the next slab fills the tail of the LZ kernel), and they shorten the two exposed
// ends of the pipeline: the first slab's way up and the last slab's way down.
constexpr uint64_t kHostSlab = 64ull << 20;
constexpr uint64_t kHostFirstSlab = 32ull << 20;
constexpr uint32_t kMaxHostSlabs = 4096;
struct HostPipe {
    zk_ctx *c;
    const uint8_t *d_out;  // device stream (2 header bytes, then the body)
    uint8_t *out;          // caller's buffer
    uint64_t lead;         // bytes in front of the body (2 for a framed stream, 0 for a shard)
    uint64_t body_cap;     // the body may not grow past this (an overflowing cursor is reported, not copied)
    uint64_t *h_cursor;    // pinned: body bytes after slab k
    cudaEvent_t *in_ready, *done;
    uint32_t nslabs;
    uint32_t flushed;      // slabs whose output has been sent down
    uint64_t sent;         // body bytes sent down so far
    cudaError_t err;
};
void pipe_flush(HostPipe *p, uint32_t upto) {  // send down the output of slabs [flushed, upto)
    for (; p->flushed < upto; ++p->flushed) {
        const uint32_t k = p->flushed;
        cudaError_t e = cudaEventSynchronize(p->done[k]);
        if (e != cudaSuccess && p->err == cudaSuccess) p->err = e;
        const uint64_t cur = p->h_cursor[k] < p->body_cap ? p->h_cursor[k] : p->body_cap;
        if (cur > p->sent) {
            e = cudaMemcpyAsync(p->out + p->lead + p->sent, p->d_out + p->lead + p->sent, cur - p->sent,
                                cudaMemcpyDeviceToHost, p->c->copy_out);
            if (e != cudaSuccess && p->err == cudaSuccess) p->err = e;
            p->c->d2h_bytes += (int64_t)(cur - p->sent);
            p->sent = cur;
        }
    }
}
void pipe_before(void *user, uint32_t k) {
    HostPipe *p = static_cast<HostPipe *>(user);
    const cudaError_t e = cudaStreamWaitEvent(p->c->stream, p->in_ready[k], 0);
    if (e != cudaSuccess && p->err == cudaSuccess) p->err = e;
}
void pipe_after(void *user, uint32_t k, const uint64_t *d_cursor) {
    HostPipe *p = static_cast<HostPipe *>(user);
    cudaError_t e = cudaMemcpyAsync(&p->h_cursor[k], d_cursor, 8, cudaMemcpyDeviceToHost, p->c->stream);
    if (e == cudaSuccess) e = cudaEventRecord(p->done[k], p->c->stream);
    if (e != cudaSuccess && p->err == cudaSuccess) p->err = e;
    if (k >= 1) pipe_flush(p, k);  // slab k is queued behind it: the device stays busy while the host waits here
}
}  // namespace

// Host buffers -> device -> host, framed (a whole zlib stream) or as a shard body primed with `hist`.
static int host_deflate(zk_ctx *c, const uint8_t *hist, uint32_t hist_len, const uint8_t *in, uint64_t n, int level,
                        int strategy, uint8_t *out, uint64_t cap, bool framed, uint64_t *out_len, uint32_t *adler,
                        uint32_t *crc = nullptr) {
    ZK_CUDA(cudaSetDevice(c->device));
    auto_register(c, in, n);
    auto_register(c, out, cap);
    const uint64_t bound = zk_compress_bound(n);
    const uint64_t dcap = cap < bound ? cap : bound;
    const size_t in_room = ((size_t)n + 255) / 256 * 256;
    const size_t out_room = ((size_t)dcap + 255) / 256 * 256 + 256;
    Staging *s = staging_for(c, 256 + (size_t)kWindow + in_room + out_room + crc_partials_bytes(n) + 256);
    if (!s) return ZK_MEM_ERROR;
    uint64_t *d_result = reinterpret_cast<uint64_t *>(s->base);
    uint8_t *d_hist = s->base + 256;
    uint8_t *d_in = d_hist + kWindow;
    uint8_t *d_out = d_in + in_room;
    uint32_t *d_crc = reinterpret_cast<uint32_t *>(s->base + 128);  // second half of the result record
    void *d_crc_partials = d_out + out_room;
    const uint64_t lead = framed ? 2 : 0;  // bytes in front of the body in d_out
    if (hist_len) ZK_CUDA(cudaMemcpyAsync(d_hist, hist, hist_len, cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += (int64_t)hist_len;
    const uint64_t slab_bytes = c->opt_host_slab ? (uint64_t)c->opt_host_slab : kHostSlab;
    const uint64_t first_bytes = tmin<uint64_t>(c->opt_host_first_slab ? (uint64_t)c->opt_host_first_slab : kHostFirstSlab, slab_bytes);
    const uint64_t nslabs64 = n <= first_bytes ? 1 : 1 + (n - first_bytes + slab_bytes - 1) / slab_bytes;
    const uint32_t nslabs = (uint32_t)tmin<uint64_t>(nslabs64, kMaxHostSlabs + 1);
    const bool piped = level > 0 && nslabs >= 2 && nslabs <= kMaxHostSlabs && dcap >= 8;
    DeflateJob job{d_in, n, hist_len ? d_hist : nullptr, hist_len, level, strategy, d_out, dcap, framed, d_result};
    if (!piped) {
        if (n) ZK_CUDA(cudaMemcpyAsync(d_in, in, n, cudaMemcpyHostToDevice, c->stream));
        c->h2d_bytes += (int64_t)n;
        int rc = run_deflate(c, job);
        if (rc == ZK_OK && crc) rc = launch_crc32(c, 0u, d_in, n, d_crc_partials, d_crc);
        if (rc != ZK_OK) return rc;
        uint64_t res[4];
        ZK_CUDA(cudaMemcpyAsync(res, d_result, sizeof res, cudaMemcpyDeviceToHost, c->stream));
        if (crc) ZK_CUDA(cudaMemcpyAsync(crc, d_crc, 4, cudaMemcpyDeviceToHost, c->stream));
        ZK_CUDA(cudaStreamSynchronize(c->stream));
        c->d2h_bytes += 32;
        const int status = (int)(int64_t)res[2];
        if (adler) *adler = (uint32_t)res[1];
        if (status != ZK_OK) return status;
        ZK_CUDA(cudaMemcpyAsync(out, d_out, res[0], cudaMemcpyDeviceToHost, c->stream));
        ZK_CUDA(cudaStreamSynchronize(c->stream));
        c->d2h_bytes += (int64_t)res[0];
        *out_len = res[0];
        return ZK_OK;
    }
    if (!c->copy_in) ZK_CUDA(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
    if (!c->copy_out) ZK_CUDA(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
    std::vector<cudaEvent_t> ev(2 * (size_t)nslabs, nullptr);
    // pinned words for the per-slab cursors: the context's small pinned block when they fit in it
    uint64_t *h_cursor = nullptr;
    const bool own_pinned = !c->pinned || (size_t)nslabs * 8 > c->pinned_cap;
    if (own_pinned)
        ZK_CUDA(cudaMallocHost(&h_cursor, (size_t)nslabs * 8));
    else
        h_cursor = reinterpret_cast<uint64_t *>(c->pinned);
    int rc = ZK_OK;
    const uint64_t frame = framed ? 2 + 2 + 4 : 0;
    HostPipe pipe{c, d_out, out, lead, dcap > frame ? dcap - frame : 0, h_cursor, ev.data(), ev.data() + nslabs, nslabs, 0, 0, cudaSuccess};
    for (auto &e : ev)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) rc = ZK_ERRNO;
    uint64_t res[4] = {0, 0, 0, 0};
    if (rc == ZK_OK) {
        // the staging buffer may still be in use by earlier work on the compute stream
        cudaEvent_t fence = pipe.done[0];
        cudaEventRecord(fence, c->stream);
        cudaStreamWaitEvent(c->copy_in, fence, 0);
        for (uint32_t k = 0; k < nslabs && pipe.err == cudaSuccess; ++k) {
            const uint64_t off = k ? first_bytes + (uint64_t)(k - 1) * slab_bytes : 0;
            const uint64_t len = tmin<uint64_t>(k ? slab_bytes : first_bytes, n - off);
            pipe.err = cudaMemcpyAsync(d_in + off, in + off, len, cudaMemcpyHostToDevice, c->copy_in);
            if (pipe.err == cudaSuccess) pipe.err = cudaEventRecord(pipe.in_ready[k], c->copy_in);
        }
        c->h2d_bytes += (int64_t)n;
        job.slab_bytes = slab_bytes;
        job.first_slab_bytes = first_bytes;
        job.before_slab = pipe_before;
        job.after_slab = pipe_after;
        job.user = &pipe;
        rc = run_deflate(c, job);
        if (rc == ZK_OK && crc) rc = launch_crc32(c, 0u, d_in, n, d_crc_partials, d_crc);  // behind every slab's kernels
        if (rc == ZK_OK && pipe.err == cudaSuccess) {
            pipe.err = cudaMemcpyAsync(res, d_result, sizeof res, cudaMemcpyDeviceToHost, c->stream);
            if (crc && pipe.err == cudaSuccess) pipe.err = cudaMemcpyAsync(crc, d_crc, 4, cudaMemcpyDeviceToHost, c->stream);
            pipe_flush(&pipe, nslabs);  // waits for the last slab; the trailer kernels are behind it
            if (pipe.err == cudaSuccess) pipe.err = cudaStreamSynchronize(c->stream);
            c->d2h_bytes += 32;
            const int status = (int)(int64_t)res[2];
            if (pipe.err == cudaSuccess && status == ZK_OK) {
                // header, and whatever follows the last slab's bytes (final block + Adler-32)
                if (lead) pipe.err = cudaMemcpyAsync(out, d_out, lead, cudaMemcpyDeviceToHost, c->copy_out);
                if (pipe.err == cudaSuccess && res[0] > lead + pipe.sent)
                    pipe.err = cudaMemcpyAsync(out + lead + pipe.sent, d_out + lead + pipe.sent, res[0] - lead - pipe.sent,
                                               cudaMemcpyDeviceToHost, c->copy_out);
                c->d2h_bytes += (int64_t)(res[0] - pipe.sent);
            }
        }
        cudaStreamSynchronize(c->copy_in);
        cudaStreamSynchronize(c->copy_out);
        cudaStreamSynchronize(c->stream);
    }
    for (auto &e : ev)
        if (e) cudaEventDestroy(e);
    if (own_pinned) cudaFreeHost(h_cursor);
    if (rc != ZK_OK) return rc;
    if (pipe.err != cudaSuccess) {
        set_last_error("host deflate (pipelined)", pipe.err);
        return ZK_ERRNO;
    }
    const int status = (int)(int64_t)res[2];
    if (adler) *adler = (uint32_t)res[1];
    if (status != ZK_OK) return status;
    *out_len = res[0];
    return ZK_OK;
}

int zk_deflate_zlib(zk_ctx *c, const uint8_t *in, uint64_t n, int level, int strategy, uint8_t *out, uint64_t cap,
                    uint64_t *out_len, uint32_t *adler) {
    if (out_len) *out_len = 0;
    if (!c || !out_len || (!in && n) || (!out && cap)) return ZK_STREAM_ERROR;
    if (level > 9) return ZK_STREAM_ERROR;  // DS:323-325: nothing written
    if (strategy < 0 || strategy > ZK_HUFFMAN_ONLY) return ZK_STREAM_ERROR;
    return host_deflate(c, nullptr, 0, in, n, level, strategy, out, cap, true, out_len, adler);
}

int zk_deflate_shard(zk_ctx *c, const uint8_t *hist, uint32_t hist_len, const uint8_t *in, uint64_t n, int level,
                     int strategy, uint8_t *out_body, uint64_t cap, uint64_t *body_len, uint32_t *adler) {
    if (body_len) *body_len = 0;
    if (!c || !body_len || (!in && n) || (!out_body && cap) || hist_len > (uint32_t)kWindow || (!hist && hist_len))
        return ZK_STREAM_ERROR;
    if (level > 9) return ZK_STREAM_ERROR;
    if (strategy < 0 || strategy > ZK_HUFFMAN_ONLY) return ZK_STREAM_ERROR;
    return host_deflate(c, hist, hist_len, in, n, level, strategy, out_body, cap, false, body_len, adler);
}

// ---- pipelined host-buffer inflate ---------------------------------------------------------------------------------
// InflateStream.inflateZlib as a caller with host buffers sees it (IS:376-438).  In one pass the call is the sum of its
// three legs: compressed bytes up, decode, output down (8 + 14 + 19.5 ms for 1 GiB in round 1).  A stream of this engine
// has a sync marker after every block, so it can be cut into slabs at markers and every slab decoded as a shard body
// (zk_inflate_shard_prepare_device: references into the output before the slab stay symbolic) on its own context, stream
// and scratch, while the slabs behind it are still arriving; a slab is finished (zk_inflate_shard_finish_device) as soon
// as the slab before it has been: its entry window is then simply the 32 KiB in front of it in the output buffer, and
// its bytes go down to the host while the next slabs decode.  The cut points are found on the host with a short search
// for 00 00 FF FF behind evenly spaced offsets; a false marker makes that slab's decode fail its validation, and then --
// as for anything else that is not plainly a stream of this shape -- the call falls back to the one-pass path, which
// owns the reference's return codes.  Returns kNotPiped when it did not handle the call.
//
// MEASURED (profiles/inflate_pipeline_r2.md, 1 GiB mixed corpus, one B200): the pipeline is correct and bit-exact but SLOWER
// than one pass -- 50.7 ms with four slabs, 89 ms with eleven, against 41.2 ms -- because a slab's prepare costs ~12 ms
// whatever its size (the lane decoder and the expander both run at the latency of ONE 32 KiB block decoded serially) and
// the prepares of concurrent slabs overlap badly.  It is therefore opt-in ("inflate_pipe" = 1) until the decoder has
// sub-block parallelism; the one-pass path stays the default.
namespace {
constexpr int kNotPiped = -1000;
constexpr uint64_t kInflatePipeMin = 48ull << 20;   // compressed bytes from which the pipeline pays
constexpr uint64_t kInflateSlab = 40ull << 20;      // compressed bytes per slab
constexpr int kPipeWorkers = 4;

struct PipeSlab {
    uint64_t in_off = 0, in_len = 0, cap = 0;   // compressed range, output room reserved in the worker's scratch
    uint64_t out_off = 0, out_len = 0;
    uint32_t adler = 1;
    cudaEvent_t in_ready = nullptr, done = nullptr;
    bool prepared = false, finished = false;
    double t_start = 0, t_prepared = 0, t_finish_begin = 0, t_finished = 0;  // ms since the call began (ZK_PIPE_DEBUG)
};
struct PipeState {
    zk_ctx *c;
    const uint8_t *d_in;
    uint8_t *d_out, *out;
    uint64_t cap;
    std::vector<PipeSlab> slabs;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<bool> failed{false};
    std::atomic<uint32_t> next{0};
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

// one worker: takes slabs in order, prepares them, and finishes each once its predecessor has been finished
void pipe_worker(PipeState *ps, zk_ctx *w, uint8_t *d_scratch) {
    if (cudaSetDevice(ps->c->device) != cudaSuccess) {
        ps->failed = true;
        ps->cv.notify_all();
        return;
    }
    uint64_t *d_result = reinterpret_cast<uint64_t *>(d_scratch);
    uint16_t *d_tail = reinterpret_cast<uint16_t *>(d_scratch + 256);
    const uint8_t *d_zero_window = d_scratch + 256 + 65536;  // 32 KiB of zeros: the window in front of the first slab
    for (;;) {
        const uint32_t k = ps->next.fetch_add(1);
        if (k >= ps->slabs.size() || ps->failed) break;
        PipeSlab &sl = ps->slabs[k];
        sl.t_start = ps->ms();
        bool ok = cudaStreamWaitEvent(w->stream, sl.in_ready, 0) == cudaSuccess;
        if (ok) ok = zk_inflate_shard_prepare_device(w, ps->d_in + sl.in_off, sl.in_len, k ? 32768u : 0u, sl.cap, d_tail, d_result) == ZK_OK;
        uint64_t res[4] = {0, 0, 0, 0};
        if (ok) ok = cudaMemcpyAsync(res, d_result, sizeof res, cudaMemcpyDeviceToHost, w->stream) == cudaSuccess &&
                     cudaStreamSynchronize(w->stream) == cudaSuccess;
        // a shard body is good when every block was decoded and it ended exactly at its end without a final block
        if (ok) ok = (int)(int64_t)res[2] == ZK_OK && res[1] == sl.in_len && res[0] <= sl.cap && w->shard.valid;
        uint64_t out_off = 0;
        {
            std::unique_lock<std::mutex> lk(ps->mu);
            if (!ok) ps->failed = true;
            sl.out_len = res[0];
            sl.prepared = true;
            sl.t_prepared = ps->ms();
            ps->cv.notify_all();
            // the slab before this one must have been finished: its output is this slab's entry window
            ps->cv.wait(lk, [&] { return ps->failed || k == 0 || ps->slabs[k - 1].finished; });
            if (ps->failed) break;
            out_off = k ? ps->slabs[k - 1].out_off + ps->slabs[k - 1].out_len : 0;
            sl.out_off = out_off;
            if (out_off + sl.out_len > ps->cap) {
                ps->failed = true;
                ps->cv.notify_all();
                break;
            }
        }
        sl.t_finish_begin = ps->ms();
        ok = k == 0 || cudaStreamWaitEvent(w->stream, ps->slabs[k - 1].done, 0) == cudaSuccess;
        // (the first slab of a zlib stream has nothing in front of it; a later one the 32 KiB before its output)
        const uint8_t *entry = k ? ps->d_out + out_off - 32768 : d_zero_window;
        if (k && out_off < 32768) ok = false;  // cannot be: every slab before it holds whole 32 KiB blocks
        if (ok) ok = zk_inflate_shard_finish_device(w, entry, ps->d_out + out_off, d_result) == ZK_OK;
        if (ok) ok = cudaEventRecord(sl.done, w->stream) == cudaSuccess;
        if (ok) ok = cudaStreamWaitEvent(ps->c->copy_out, sl.done, 0) == cudaSuccess;
        if (ok && sl.out_len)
            ok = cudaMemcpyAsync(ps->out + out_off, ps->d_out + out_off, sl.out_len, cudaMemcpyDeviceToHost, ps->c->copy_out) == cudaSuccess;
        if (ok) ok = cudaMemcpyAsync(res, d_result, sizeof res, cudaMemcpyDeviceToHost, w->stream) == cudaSuccess &&
                     cudaStreamSynchronize(w->stream) == cudaSuccess;
        if (ok) ok = (int)(int64_t)res[2] == ZK_OK || (int)(int64_t)res[2] == ZK_STREAM_END;
        {
            std::lock_guard<std::mutex> lk(ps->mu);
            if (!ok) ps->failed = true;
            sl.adler = (uint32_t)res[3];
            sl.finished = true;
            sl.t_finished = ps->ms();
        }
        ps->cv.notify_all();
        if (!ok) break;
    }
}

int host_inflate_piped(zk_ctx *c, const uint8_t *in, uint64_t m, uint8_t *out, uint64_t cap, uint64_t *out_len, uint64_t *consumed) {
    if (!c->opt_inflate_pipe || m < kInflatePipeMin || cap < (1u << 20)) return kNotPiped;
    // a zlib stream without a dictionary whose last block is the empty fixed block this engine (and pigz) ends with
    if (((in[0] << 8) | in[1]) % 31 != 0 || (in[0] & 15) != 8 || (in[1] & 0x20)) return kNotPiped;
    if (in[m - 6] != 0x03 || in[m - 5] != 0x00 || memcmp(in + m - 10, "\0\0\xFF\xFF", 4) != 0) return kNotPiped;
    const uint64_t body_begin = 2, body_end = m - 6;
    const uint64_t slab_bytes = c->opt_inflate_slab ? (uint64_t)c->opt_inflate_slab : kInflateSlab;
    // cut points: behind the first marker after every multiple of slab_bytes
    std::vector<uint64_t> cuts{body_begin};
    for (uint64_t target = body_begin + slab_bytes; target + (1u << 20) < body_end; target += slab_bytes) {
        const uint8_t *from = in + target, *end = in + tmin<uint64_t>(target + (1u << 20), body_end);
        const uint8_t *hit = nullptr;
        for (const uint8_t *p = from; p + 4 <= end;) {
            p = static_cast<const uint8_t *>(memchr(p, 0xFF, (size_t)(end - p)));
            if (!p) break;
            if (p - in >= 2 && p + 1 < end && p[1] == 0xFF && p[-1] == 0 && p[-2] == 0) {
                hit = p + 2;
                break;
            }
            ++p;
        }
        if (!hit) return kNotPiped;
        cuts.push_back((uint64_t)(hit - in));
    }
    const uint32_t nslabs = (uint32_t)cuts.size();
    if (nslabs < 2) return kNotPiped;
    if (cudaSetDevice(c->device) != cudaSuccess) return kNotPiped;
    int workers = (int)tmin<uint32_t>(nslabs, kPipeWorkers);
    if (const char *e = getenv("ZK_PIPE_WORKERS")) workers = tmax(1, tmin(workers, atoi(e)));  // tuning aid
    for (int w = 0; w < workers; ++w)
        if (!c->helpers[w]) {
            if (!(c->helpers[w] = zk_ctx_create(c->device))) return kNotPiped;
            c->helpers[w]->opt_lane_pack = 1;
        }
    const size_t in_room = ((size_t)m + 255) / 256 * 256;
    const size_t per_worker = 256 + 65536 + 32768 + 256;
    Staging *st = staging_for(c, 256 + in_room + (size_t)cap + 256 + (size_t)workers * per_worker);
    if (!st) return kNotPiped;
    uint8_t *d_in = st->base + 256, *d_out = d_in + in_room, *d_work = d_out + ((size_t)cap + 255) / 256 * 256;
    if (!c->copy_in && cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking) != cudaSuccess) return kNotPiped;
    if (!c->copy_out && cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking) != cudaSuccess) return kNotPiped;
    PipeState ps;
    ps.c = c;
    ps.d_in = d_in;
    ps.d_out = d_out;
    ps.out = out;
    ps.cap = cap;
    ps.slabs.resize(nslabs);
    bool ok = true;
    // output room per slab in its worker's scratch: its share of the caller's buffer by compressed size, times four (a
    // slab that expands more than that sends the call to the one-pass path)
    for (uint32_t k = 0; k < nslabs; ++k) {
        PipeSlab &sl = ps.slabs[k];
        sl.in_off = cuts[k];
        sl.in_len = (k + 1 < nslabs ? cuts[k + 1] : body_end) - cuts[k];
        const double share = (double)sl.in_len / (double)(body_end - body_begin);
        sl.cap = tmin<uint64_t>(cap, (uint64_t)(share * 4.0 * (double)cap) + (1u << 20));
        ok = ok && cudaEventCreateWithFlags(&sl.in_ready, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming) == cudaSuccess;
    }
    auto cleanup = [&] {
        for (PipeSlab &sl : ps.slabs) {
            if (sl.in_ready) cudaEventDestroy(sl.in_ready);
            if (sl.done) cudaEventDestroy(sl.done);
        }
    };
    if (!ok) {
        cleanup();
        return kNotPiped;
    }
    // the staging buffer may still be in use by earlier work on the compute stream
    cudaEventRecord(ps.slabs[0].done, c->stream);
    cudaStreamWaitEvent(c->copy_in, ps.slabs[0].done, 0);
    cudaStreamWaitEvent(c->copy_out, ps.slabs[0].done, 0);
    ok = cudaMemsetAsync(d_work, 0, (size_t)workers * per_worker, c->copy_in) == cudaSuccess;
    for (uint32_t k = 0; k < nslabs && ok; ++k) {
        PipeSlab &sl = ps.slabs[k];
        // (slab 0 also carries the two header bytes in front of it, so that offsets in d_in equal offsets in `in`)
        const uint64_t from = k ? sl.in_off : 0, len = sl.in_off + sl.in_len - from + (k + 1 == nslabs ? 6 : 0);
        ok = cudaMemcpyAsync(d_in + from, in + from, len, cudaMemcpyHostToDevice, c->copy_in) == cudaSuccess &&
             cudaEventRecord(sl.in_ready, c->copy_in) == cudaSuccess;
    }
    c->h2d_bytes += (int64_t)m;
    if (ok) {
        std::vector<std::thread> threads;
        for (int w = 0; w < workers; ++w) threads.emplace_back(pipe_worker, &ps, c->helpers[w], d_work + (size_t)w * per_worker);
        for (std::thread &t : threads) t.join();
    }
    cudaStreamSynchronize(c->copy_in);
    cudaStreamSynchronize(c->copy_out);
    for (int w = 0; w < workers; ++w) {
        cudaStreamSynchronize(c->helpers[w]->stream);
        c->launches += c->helpers[w]->launches;
        c->helpers[w]->launches = 0;
    }
    cleanup();
    if (getenv("ZK_PIPE_DEBUG")) {
        fprintf(stderr, "[pipe] %u slabs, %d workers, total %.2f ms\n", nslabs, workers, ps.ms());
        for (uint32_t k = 0; k < nslabs; ++k)
            fprintf(stderr, "[pipe] slab %2u in %6.1f MiB out %6.1f MiB  start %6.2f prepared %6.2f finish %6.2f .. %6.2f\n", k,
                    ps.slabs[k].in_len / 1048576.0, ps.slabs[k].out_len / 1048576.0, ps.slabs[k].t_start, ps.slabs[k].t_prepared,
                    ps.slabs[k].t_finish_begin, ps.slabs[k].t_finished);
    }
    if (!ok || ps.failed) {
        c->pipe_fallbacks++;
        return kNotPiped;
    }
    uint64_t total = 0;
    uint32_t adler = 1;
    for (const PipeSlab &sl : ps.slabs) {
        total += sl.out_len;
        adler = adler32_combine(adler, sl.adler, sl.out_len);
    }
    c->d2h_bytes += (int64_t)total;
    c->pipe_slabs += nslabs;
    const uint32_t want = ((uint32_t)in[m - 4] << 24) | ((uint32_t)in[m - 3] << 16) | ((uint32_t)in[m - 2] << 8) | in[m - 1];
    *out_len = total;
    if (consumed) *consumed = m;
    return adler == want ? ZK_STREAM_END : ZK_DATA_ERROR;  // IS:419-424
}
}  // namespace

static int host_inflate(zk_ctx *c, const uint8_t *in, uint64_t m, const uint8_t *dict, uint32_t dict_len, uint8_t *out,
                        uint64_t cap, uint64_t *out_len, uint64_t *consumed, int container = 0, uint32_t body_begin = 0,
                        uint32_t *check = nullptr, bool quick_incomplete = false, uint32_t begin_bit = 0, bool prefix = false) {
    if (out_len) *out_len = 0;
    if (consumed) *consumed = 0;
    if (!c || !out_len || (!in && m) || (!out && cap) || (!dict && dict_len)) return ZK_STREAM_ERROR;
    ZK_CUDA(cudaSetDevice(c->device));
    auto_register(c, in, m);
    auto_register(c, out, cap);
    if (container == 0 && !dict && m >= 16) {
        const int prc = host_inflate_piped(c, in, m, out, cap, out_len, consumed);
        if (prc != kNotPiped) {
            if (check) *check = ((uint32_t)in[m - 4] << 24) | ((uint32_t)in[m - 3] << 16) | ((uint32_t)in[m - 2] << 8) | in[m - 1];
            return prc;
        }
        *out_len = 0;
        if (consumed) *consumed = 0;
    }
    uint32_t dict_id = 0;
    if (dict && container == 0) dict_id = dict_len ? zk_adler32(c, 1u, dict, dict_len) : 1u;  // DICTID = Adler-32 of the dictionary
    const size_t in_room = ((size_t)m + 255) / 256 * 256, dict_room = ((size_t)dict_len + 255) / 256 * 256 + 256;
    Staging *s = staging_for(c, 256 + dict_room + in_room + (size_t)cap + 256);
    if (!s) return ZK_MEM_ERROR;
    uint64_t *d_result = reinterpret_cast<uint64_t *>(s->base);
    uint8_t *d_dict = s->base + 256;
    uint8_t *d_in = d_dict + dict_room;
    uint8_t *d_out = d_in + in_room;
    if (dict_len) ZK_CUDA(cudaMemcpyAsync(d_dict, dict, dict_len, cudaMemcpyHostToDevice, c->stream));
    if (m) ZK_CUDA(cudaMemcpyAsync(d_in, in, m, cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += (int64_t)m + dict_len;
    InflateJob job{d_in, m, d_out, cap, d_result};
    if (dict) {
        job.d_dict = d_dict;
        job.dict_len = dict_len;
        job.dict_id = dict_id;
    }
    job.container = container;
    job.body_begin = body_begin;
    job.quick_incomplete = quick_incomplete;
    job.begin_bit = begin_bit;
    job.prefix = prefix;
    int rc = run_inflate(c, job);
    if (rc != ZK_OK) return rc;
    uint64_t res[4];
    ZK_CUDA(cudaMemcpyAsync(res, d_result, sizeof res, cudaMemcpyDeviceToHost, c->stream));
    ZK_CUDA(cudaStreamSynchronize(c->stream));
    c->d2h_bytes += 32;
    const int status = (int)(int64_t)res[2];
    if (check) *check = (uint32_t)res[3];
    if (res[0] && res[0] <= cap) {
        ZK_CUDA(cudaMemcpyAsync(out, d_out, res[0], cudaMemcpyDeviceToHost, c->stream));
        ZK_CUDA(cudaStreamSynchronize(c->stream));
        c->d2h_bytes += (int64_t)res[0];
    }
    *out_len = res[0] <= cap ? res[0] : 0;
    if (consumed) *consumed = res[1];
    return status;
}

int zk_inflate_zlib(zk_ctx *c, const uint8_t *in, uint64_t m, uint8_t *out, uint64_t cap, uint64_t *out_len,
                    uint64_t *consumed) {
    return host_inflate(c, in, m, nullptr, 0, out, cap, out_len, consumed);
}

int zk_inflate_zlib_dict(zk_ctx *c, const uint8_t *in, uint64_t m, const uint8_t *dict, uint32_t dict_len, uint8_t *out,
                         uint64_t cap, uint64_t *out_len, uint64_t *consumed) {
    if (!dict) return ZK_STREAM_ERROR;
    return host_inflate(c, in, m, dict, dict_len, out, cap, out_len, consumed);
}

int zk_deflate_zlib_dict(zk_ctx *c, const uint8_t *dict, uint32_t dict_len, const uint8_t *in, uint64_t n, int level,
                         int strategy, uint8_t *out, uint64_t cap, uint64_t *out_len, uint32_t *adler) {
    // zlib.h deflateSetDictionary: FDICT in the header, DICTID = Adler-32 of the dictionary, matching primed with its last
    // 32 KiB; the trailer covers the data only
    if (out_len) *out_len = 0;
    if (!c || !out_len || !dict || (!in && n) || !out) return ZK_STREAM_ERROR;
    if (level > 9 || strategy < 0 || strategy > ZK_HUFFMAN_ONLY) return ZK_STREAM_ERROR;
    if (cap < 12) return ZK_BUF_ERROR;
    const uint32_t id = dict_len ? zk_adler32(c, 1u, dict, dict_len) : 1u;
    const uint32_t keep = dict_len > (uint32_t)kWindow ? (uint32_t)kWindow : dict_len;
    uint64_t body = 0;
    uint32_t a = 1;
    int rc = zk_deflate_shard(c, dict + (dict_len - keep), keep, in, n, level, strategy, out + 6, cap - 12, &body, &a);
    if (rc != ZK_OK) return rc;
    uint8_t h[2];
    zk_zlib_header(level, h);
    h[1] = (uint8_t)((h[1] & 0xC0) | 0x20);
    h[1] = (uint8_t)(h[1] + (31 - ((h[0] << 8) | h[1]) % 31) % 31);
    out[0] = h[0], out[1] = h[1];
    out[2] = (uint8_t)(id >> 24), out[3] = (uint8_t)(id >> 16), out[4] = (uint8_t)(id >> 8), out[5] = (uint8_t)id;
    zk_zlib_tail(a, out + 6 + body);
    *out_len = 6 + body + 6;
    if (adler) *adler = a;
    return ZK_OK;
}

// ---- CRC-32 and the gzip / raw containers (RFC 1952 / RFC 1951) -------------------------------
uint32_t zk_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) { return crc32_combine(crc1, crc2, len2); }

int zk_crc32_device(zk_ctx *c, uint32_t crc, const uint8_t *d_buf, uint64_t len, uint32_t *d_out) {
    if (!c || !d_out || (!d_buf && len)) return ZK_STREAM_ERROR;
    ZK_CUDA(cudaSetDevice(c->device));
    c->arena.reset();
    c->shard.valid = false;  // the scratch of a pending sharded inflate is gone
    if (c->arena.reserve(crc_partials_bytes(len) + 1024) != ZK_OK) return ZK_MEM_ERROR;
    void *d_partials = c->arena.take(crc_partials_bytes(len));
    return launch_crc32(c, crc, d_buf, len, d_partials, d_out);
}

uint32_t zk_crc32(zk_ctx *c, uint32_t crc, const uint8_t *buf, uint64_t len) {
    if (buf == nullptr) return 0u;  // zlib.h: crc32(0, Z_NULL, 0) is the initial value
    if (!c || len == 0) return crc;
    if (cudaSetDevice(c->device) != cudaSuccess) return 0u;
    Staging *s = staging_for(c, (size_t)len + 512);
    if (!s) return 0u;
    uint32_t *d_out = reinterpret_cast<uint32_t *>(s->base);
    uint8_t *d_buf = s->base + 256;
    if (cudaMemcpyAsync(d_buf, buf, len, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return 0u;
    c->h2d_bytes += (int64_t)len;
    if (zk_crc32_device(c, crc, d_buf, len, d_out) != ZK_OK) return 0u;
    uint32_t out = 0;
    if (cudaMemcpyAsync(&out, d_out, 4, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return 0u;
    cudaStreamSynchronize(c->stream);
    c->d2h_bytes += 4;
    return out;
}

int zk_deflate_raw(zk_ctx *c, const uint8_t *in, uint64_t n, int level, int strategy, uint8_t *out, uint64_t cap,
                   uint64_t *out_len) {
    // a complete raw deflate stream (RFC 1951): the shard body and the final empty fixed block (level <= 0: the body is
    // stored blocks, as in the zlib container, DS:307-308)
    if (out_len) *out_len = 0;
    if (!c || !out_len || (!in && n) || !out) return ZK_STREAM_ERROR;
    if (level > 9 || strategy < 0 || strategy > ZK_HUFFMAN_ONLY) return ZK_STREAM_ERROR;
    if (cap < 2) return ZK_BUF_ERROR;
    uint64_t body = 0;
    int rc = host_deflate(c, nullptr, 0, in, n, level, strategy, out, cap - 2, false, &body, nullptr);
    if (rc != ZK_OK) return rc;
    out[body] = 0x03, out[body + 1] = 0x00;
    *out_len = body + 2;
    return ZK_OK;
}

int zk_deflate_gzip(zk_ctx *c, const uint8_t *in, uint64_t n, int level, int strategy, uint8_t *out, uint64_t cap,
                    uint64_t *out_len, uint32_t *crc) {
    // one gzip member (RFC 1952): 10-byte header without optional fields, raw deflate, CRC-32 and ISIZE (little-endian)
    if (out_len) *out_len = 0;
    if (!c || !out_len || (!in && n) || !out) return ZK_STREAM_ERROR;
    if (level > 9 || strategy < 0 || strategy > ZK_HUFFMAN_ONLY) return ZK_STREAM_ERROR;
    if (cap < 20) return ZK_BUF_ERROR;
    const uint8_t head[10] = {0x1F, 0x8B, 8, 0, 0, 0, 0, 0, (uint8_t)(level >= 9 ? 2 : level <= 1 ? 4 : 0), 3 /* Unix */};
    uint64_t body = 0;
    uint32_t cr = 0;
    int rc = host_deflate(c, nullptr, 0, in, n, level, strategy, out + 10, cap - 20, false, &body, nullptr, &cr);
    if (rc != ZK_OK) return rc;
    memcpy(out, head, 10);
    uint8_t *t = out + 10 + body;
    t[0] = 0x03, t[1] = 0x00;
    for (int i = 0; i < 4; ++i) {
        t[2 + i] = (uint8_t)(cr >> (8 * i));
        t[6 + i] = (uint8_t)((uint32_t)n >> (8 * i));
    }
    *out_len = 10 + body + 10;
    if (crc) *crc = cr;
    return ZK_OK;
}

int zk_inflate_raw(zk_ctx *c, const uint8_t *in, uint64_t m, uint8_t *out, uint64_t cap, uint64_t *out_len,
                   uint64_t *consumed) {
    return host_inflate(c, in, m, nullptr, 0, out, cap, out_len, consumed, 1, 0);
}

// RFC 1952 member header -> offset of the deflate data, or 0 (bad header) / ~0u (truncated)
static uint32_t gzip_header_len(const uint8_t *p, uint64_t m) {
    if (m < 10) return ~0u;
    if (p[0] != 0x1F || p[1] != 0x8B || p[2] != 8 || (p[3] & 0xE0)) return 0;
    const int flg = p[3];
    uint64_t at = 10;
    if (flg & 4) {  // FEXTRA
        if (at + 2 > m) return ~0u;
        at += 2 + ((uint64_t)p[at] | ((uint64_t)p[at + 1] << 8));
    }
    for (int bit = 8; bit <= 16; bit <<= 1)  // FNAME, FCOMMENT: zero-terminated
        if (flg & bit) {
            while (at < m && p[at]) ++at;
            ++at;
        }
    if (flg & 2) at += 2;  // FHCRC
    if (at > m) return ~0u;
    return at > 0x7FFFFFFFull ? 0 : (uint32_t)at;
}

int zk_inflate_gzip(zk_ctx *c, const uint8_t *in, uint64_t m, uint8_t *out, uint64_t cap, uint64_t *out_len,
                    uint64_t *consumed) {
    // all members of the file (RFC 1952 2.2: a gzip file is a series of members), outputs concatenated
    if (out_len) *out_len = 0;
    if (consumed) *consumed = 0;
    if (!c || !out_len || (!in && m) || (!out && cap)) return ZK_STREAM_ERROR;
    uint64_t at = 0, made = 0;
    int members = 0;
    while (at < m) {
        const uint32_t h = gzip_header_len(in + at, m - at);
        if (h == 0) {
            if (members) break;  // trailing bytes that are not a member are left unread
            return ZK_DATA_ERROR;
        }
        if (h == ~0u) return members ? ZK_STREAM_END : ZK_BUF_ERROR;
        uint64_t got = 0, used = 0;
        const int rc = host_inflate(c, in + at, m - at, nullptr, 0, out + made, cap - made, &got, &used, 2, h);
        made += got;
        *out_len = made;
        if (rc != ZK_STREAM_END) {
            if (consumed) *consumed = at + used;
            return rc;
        }
        at += used;
        ++members;
        if (consumed) *consumed = at;
    }
    return members ? ZK_STREAM_END : ZK_BUF_ERROR;
}

// ---- streaming handles (docs/API.md:205-281) ------------------------------------------------------------------
// deflate: input is gathered into slabs; a slab is compressed as a headerless body primed with the previous
// 32 KiB (the same mechanism as a shard), so output produced for different slabs concatenates byte-wise.
// inflate: the compressed input is gathered until the stream can be decoded as a whole.
struct zk_stream {
    zk_ctx *ctx;
    int mode;  // 0 deflate, 1 inflate
    int level, strategy;
    bool header_done, finished;
    uint64_t total_in, total_out;
    uint32_t adler;
    std::vector<uint8_t> pending_in;   // input not yet compressed / decoded
    std::vector<uint8_t> history;      // last 32 KiB of compressed-so-far input (deflate) / dictionary
    std::vector<uint8_t> pending_out;  // produced, not yet handed to the caller
    size_t out_pos;
    int container = 0;                 // 0 zlib, 1 raw deflate (windowBits < 0), 2 gzip (windowBits + 16)
    uint32_t crc = 0;                  // gzip deflate: CRC-32 of the input so far
    bool have_dict = false;            // deflate: emit FDICT + DICTID; inflate: a dictionary has been supplied
    uint32_t dict_id = 0;
    std::vector<uint8_t> dict;         // inflate: the supplied dictionary
    // inflate state machine: what is carried from call to call (SURVEY.md section 5 "checkpoint / resume")
    std::vector<uint8_t> window;       // the last <= 32 KiB of output: entry window of the next segment
    size_t scan_pos = 0;               // pending_in[0, scan_pos) has been searched for sync markers
    size_t marker_end = 0;             // end of the last marker found in pending_in (0: none)
    size_t tail_attempt = 0;           // size of pending_in when the end of the stream was last tried and found incomplete
    bool seen_marker = false;          // a sync marker has been decoded up to: the stream is cut into short blocks
    bool plain_mode = false;           // a marker turned out not to be a block boundary: gather, try at geometric sizes
    bool body_done = false;            // the final block has been decoded; the trailer may still be missing
    uint64_t decoded = 0;              // bytes decoded so far (total_out counts what the caller has taken)
    uint32_t bit_off = 0;              // bits of pending_in[0] that belong to blocks already decoded (after a prefix decode)
    size_t prefix_attempt = 0;         // size of pending_in when a prefix decode last made no progress
    int sticky = ZK_OK;                // an error, once met, is the answer of every later call
};

static const size_t kStreamSlab = 8u << 20;

zk_stream *zk_deflate_init(zk_ctx *ctx, int level, int strategy) {
    if (!ctx || level > 9 || strategy < 0 || strategy > ZK_HUFFMAN_ONLY) return nullptr;
    zk_stream *s = new (std::nothrow) zk_stream();
    if (!s) return nullptr;
    s->ctx = ctx;
    s->mode = 0;
    s->level = level;
    s->strategy = strategy;
    s->header_done = s->finished = false;
    s->total_in = s->total_out = 0;
    s->adler = 1;
    s->out_pos = 0;
    return s;
}

// compress pending_in as one headerless body and append it to pending_out
static int stream_flush_slab(zk_stream *s) {
    zk_ctx *c = s->ctx;
    const size_t n = s->pending_in.size();
    if (n == 0) return ZK_OK;
    ZK_CUDA(cudaSetDevice(c->device));
    const uint64_t bound = zk_compress_bound(n);
    const size_t in_room = (n + 255) / 256 * 256;
    Staging *st = staging_for(c, 256 + 32768 + in_room + (size_t)bound + 256);
    if (!st) return ZK_MEM_ERROR;
    uint64_t *d_meta = reinterpret_cast<uint64_t *>(st->base);
    uint8_t *d_hist = st->base + 256;
    uint8_t *d_in = d_hist + 32768;
    uint8_t *d_out = d_in + in_room;
    const uint32_t hl = (uint32_t)s->history.size();
    if (hl) ZK_CUDA(cudaMemcpyAsync(d_hist, s->history.data(), hl, cudaMemcpyHostToDevice, c->stream));
    ZK_CUDA(cudaMemcpyAsync(d_in, s->pending_in.data(), n, cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += (int64_t)(n + hl);
    int rc = zk_deflate_shard_device(c, d_hist, hl, d_in, n, s->level, s->strategy, d_out, bound, d_meta);
    if (rc != ZK_OK) return rc;
    uint64_t meta[4];
    ZK_CUDA(cudaMemcpyAsync(meta, d_meta, sizeof meta, cudaMemcpyDeviceToHost, c->stream));
    ZK_CUDA(cudaStreamSynchronize(c->stream));
    if ((int)(int64_t)meta[2] != ZK_OK) return (int)(int64_t)meta[2];
    const size_t at = s->pending_out.size();
    s->pending_out.resize(at + meta[0]);
    ZK_CUDA(cudaMemcpyAsync(s->pending_out.data() + at, d_out, meta[0], cudaMemcpyDeviceToHost, c->stream));
    ZK_CUDA(cudaStreamSynchronize(c->stream));
    c->d2h_bytes += (int64_t)meta[0] + 32;
    s->adler = adler32_combine(s->adler, (uint32_t)meta[1], n);
    if (s->container == 2) s->crc = zk_crc32(c, s->crc, s->pending_in.data(), n);
    // keep the last 32 KiB as the next slab's dictionary
    std::vector<uint8_t> h;
    if (n >= 32768) h.assign(s->pending_in.end() - 32768, s->pending_in.end());
    else {
        const size_t keep = s->history.size() + n > 32768 ? 32768 - n : s->history.size();
        h.assign(s->history.end() - keep, s->history.end());
        h.insert(h.end(), s->pending_in.begin(), s->pending_in.end());
    }
    s->history.swap(h);
    s->pending_in.clear();
    return ZK_OK;
}

static int zk_deflate_impl(zk_stream *s, const uint8_t *in, uint64_t *in_len, uint8_t *out, uint64_t *out_len, int flush);
// The header promises "never throws": the streaming calls grow std::vectors, so an allocation failure becomes ZK_MEM_ERROR.
int zk_deflate(zk_stream *s, const uint8_t *in, uint64_t *in_len, uint8_t *out, uint64_t *out_len, int flush) {
    try {
        return zk_deflate_impl(s, in, in_len, out, out_len, flush);
    } catch (const std::bad_alloc &) {
        return ZK_MEM_ERROR;
    }
}
static int zk_deflate_impl(zk_stream *s, const uint8_t *in, uint64_t *in_len, uint8_t *out, uint64_t *out_len, int flush) {
    if (!s || s->mode != 0 || !in_len || !out_len || flush < 0 || flush > ZK_FINISH) return ZK_STREAM_ERROR;
    const uint64_t avail_in = *in_len, avail_out = *out_len;
    *in_len = 0;
    *out_len = 0;
    if (s->finished && avail_in) return ZK_STREAM_ERROR;
    if (!s->header_done) {
        uint8_t h[10] = {0x1F, 0x8B, 8, 0, 0, 0, 0, 0, (uint8_t)(s->level >= 9 ? 2 : s->level <= 1 ? 4 : 0), 3};
        if (s->container == 0) zk_zlib_header(s->level, h);
        if (s->container == 0 && s->have_dict) {  // zlib.h deflateSetDictionary: FDICT, then the dictionary's Adler-32
            h[1] = (uint8_t)((h[1] & 0xC0) | 0x20);
            h[1] = (uint8_t)(h[1] + (31 - ((h[0] << 8) | h[1]) % 31) % 31);
            h[2] = (uint8_t)(s->dict_id >> 24), h[3] = (uint8_t)(s->dict_id >> 16), h[4] = (uint8_t)(s->dict_id >> 8), h[5] = (uint8_t)s->dict_id;
        }
        const int hl = s->container == 2 ? 10 : s->container == 1 ? 0 : s->have_dict ? 6 : 2;
        s->pending_out.insert(s->pending_out.end(), h, h + hl);
        s->header_done = true;
    }
    if (avail_in) {
        if (!in) return ZK_STREAM_ERROR;
        s->pending_in.insert(s->pending_in.end(), in, in + avail_in);
        s->total_in += avail_in;
        *in_len = avail_in;
    }
    if (!s->finished && (s->pending_in.size() >= kStreamSlab || (flush != ZK_NO_FLUSH && !s->pending_in.empty()))) {
        int rc = stream_flush_slab(s);
        if (rc != ZK_OK) return rc;
        if (flush == ZK_FULL_FLUSH) s->history.clear();  // later data may not refer to earlier data
    }
    if (flush == ZK_FINISH && !s->finished) {
        uint8_t t[10];
        int tl = 6;
        zk_zlib_tail(s->adler, t);  // final empty block + big-endian Adler-32
        if (s->container == 1) tl = 2;
        if (s->container == 2) {
            for (int i = 0; i < 4; ++i) {
                t[2 + i] = (uint8_t)(s->crc >> (8 * i));
                t[6 + i] = (uint8_t)((uint32_t)s->total_in >> (8 * i));
            }
            tl = 10;
        }
        s->pending_out.insert(s->pending_out.end(), t, t + tl);
        s->finished = true;
    }
    const size_t have = s->pending_out.size() - s->out_pos;
    const size_t give = have < avail_out ? have : (size_t)avail_out;
    if (give) {
        if (!out) return ZK_STREAM_ERROR;
        memcpy(out, s->pending_out.data() + s->out_pos, give);
        s->out_pos += give;
        s->total_out += give;
        *out_len = give;
    }
    if (s->out_pos == s->pending_out.size()) {
        s->pending_out.clear();
        s->out_pos = 0;
        if (s->finished) return ZK_STREAM_END;
    }
    if (!*in_len && !give) return ZK_BUF_ERROR;  // zlib.h: no progress possible
    return ZK_OK;
}

int zk_deflate_reset(zk_stream *s) {
    if (!s || s->mode != 0) return ZK_STREAM_ERROR;
    s->header_done = s->finished = false;
    s->total_in = s->total_out = 0;
    s->adler = 1;
    s->crc = 0;
    s->have_dict = false;
    s->pending_in.clear();
    s->history.clear();
    s->pending_out.clear();
    s->out_pos = 0;
    return ZK_OK;
}

zk_stream *zk_deflate_init2(zk_ctx *ctx, int level, int strategy, int container) {
    if (container < 0 || container > 2) return nullptr;
    zk_stream *s = zk_deflate_init(ctx, level, strategy);
    if (s) s->container = container;
    return s;
}

zk_stream *zk_inflate_init2(zk_ctx *ctx, int container) {
    if (container < 0 || container > 2) return nullptr;
    zk_stream *s = zk_inflate_init(ctx);
    if (s) s->container = container;
    return s;
}

zk_stream *zk_deflate_copy(const zk_stream *src) {
    // zlib.h deflateCopy: an independent stream with the same state (pending input and output, history, checksums)
    if (!src || src->mode != 0) return nullptr;
    try {
        return new zk_stream(*src);
    } catch (const std::bad_alloc &) {
        return nullptr;
    }
}

int zk_deflate_params(zk_stream *s, int level, int strategy) {
    if (!s || s->mode != 0 || level > 9 || strategy < 0 || strategy > ZK_HUFFMAN_ONLY) return ZK_STREAM_ERROR;
    if (!s->pending_in.empty()) {  // zlib.h: data so far is compressed with the old parameters
        int rc = stream_flush_slab(s);
        if (rc != ZK_OK) return rc;
    }
    s->level = level;
    s->strategy = strategy;
    return ZK_OK;
}

int zk_deflate_set_dictionary(zk_stream *s, const uint8_t *dict, uint32_t len) {
    // zlib.h semantics: accepted before any input; the header then carries FDICT and the dictionary's Adler-32, and the
    // last 32 KiB of the dictionary prime the match finder (the same mechanism as a shard's history)
    if (!s || s->mode != 0 || !dict || s->total_in || s->header_done) return ZK_STREAM_ERROR;
    const uint32_t keep = len > 32768 ? 32768 : len;
    try {
        s->history.assign(dict + (len - keep), dict + len);
    } catch (const std::bad_alloc &) {
        return ZK_MEM_ERROR;
    }
    s->have_dict = true;
    s->dict_id = len ? zk_adler32(s->ctx, 1u, dict, len) : 1u;
    return ZK_OK;
}

zk_stream *zk_inflate_init(zk_ctx *ctx) {
    if (!ctx) return nullptr;
    zk_stream *s = new (std::nothrow) zk_stream();
    if (!s) return nullptr;
    s->ctx = ctx;
    s->mode = 1;
    s->level = s->strategy = 0;
    s->header_done = s->finished = false;
    s->total_in = s->total_out = 0;
    s->adler = 1;
    s->out_pos = 0;
    return s;
}

// ---- streaming inflate: a state machine over complete blocks ---------------------------------------------------------
// docs/API.md:239-308 (inflate(flush), ZInputStream.read).  Every call appends its input to the unconsumed tail and decodes
// what can be decoded NOW, once: all complete blocks up to the last sync marker that has arrived (a headerless body whose
// entry window is the last 32 KiB produced so far -- the shard mechanism), and the end of the stream (final block +
// trailer) when it looks complete.  Carried between calls: the 32 KiB window, the unconsumed tail, the running checksum
// and the totals.  A stream without markers (libz, the reference) has nothing to decode before its end: its input is
// gathered and the end is tried when the tail has doubled since the last try (or at Z_FINISH), so the GPU work stays
// linear in the stream either way.  total_in counts everything taken from the caller (as zlib does); total_out and adler
// describe what has been DECODED AND HANDED OUT, which lags behind zlib's by at most the block in flight.
namespace {

void stream_note_output(zk_stream *s, const uint8_t *p, size_t n) {
    if (n >= 32768) {
        s->window.assign(p + n - 32768, p + n);
    } else if (n) {
        const size_t keep = tmin<size_t>(s->window.size(), 32768 - n);
        std::vector<uint8_t> w(s->window.end() - (long)keep, s->window.end());
        w.insert(w.end(), p, p + n);
        s->window.swap(w);
    }
    s->decoded += n;
}

// pending_in[0, len) as a headerless body of whole blocks ending on a sync marker.  0: decoded and appended to
// pending_out; 1: it is not such a body (a false marker, damaged data, a final block inside); < 0: an error to report.
int stream_decode_marked(zk_stream *s, size_t len) {
    zk_ctx *c = s->ctx;
    if (cudaSetDevice(c->device) != cudaSuccess) return ZK_ERRNO;
    const uint32_t wl = (uint32_t)s->window.size();
    uint64_t cap = tmax<uint64_t>((uint64_t)len * 6, 1u << 20);
    for (int attempt = 0; attempt < 5; ++attempt, cap *= 8) {
        const size_t in_room = (len + 255) / 256 * 256;
        Staging *st = staging_for(c, 256 + 65536 + 32768 + in_room + (size_t)cap + 256);
        if (!st) return attempt ? 1 : ZK_MEM_ERROR;  // (a retry with eight times the room: then this is not such a body)
        uint64_t *d_result = reinterpret_cast<uint64_t *>(st->base);
        uint16_t *d_tail = reinterpret_cast<uint16_t *>(st->base + 256);
        uint8_t *d_window = st->base + 256 + 65536, *d_in = d_window + 32768, *d_out = d_in + in_room;
        ZK_CUDA(cudaMemsetAsync(d_window, 0, 32768, c->stream));
        if (wl) ZK_CUDA(cudaMemcpyAsync(d_window + 32768 - wl, s->window.data(), wl, cudaMemcpyHostToDevice, c->stream));
        ZK_CUDA(cudaMemcpyAsync(d_in, s->pending_in.data(), len, cudaMemcpyHostToDevice, c->stream));
        c->h2d_bytes += (int64_t)len + wl;
        c->stream_quick = true;  // a false marker: the body runs out of input; no need to walk it with one warp to learn that
        int rc = zk_inflate_shard_prepare_device(c, d_in, len, wl, cap, d_tail, d_result);
        c->stream_quick = false;
        if (rc != ZK_OK) return rc;
        uint64_t res[4];
        ZK_CUDA(cudaMemcpyAsync(res, d_result, sizeof res, cudaMemcpyDeviceToHost, c->stream));
        ZK_CUDA(cudaStreamSynchronize(c->stream));
        const int status = (int)(int64_t)res[2];
        // more output than reserved (nothing written, nothing consumed): once more.  A body that runs out of INPUT -- a
        // false marker, e.g. the 00 | 00 FF FF of a stored block header behind a zero byte -- reports what it consumed.
        if (status == ZK_BUF_ERROR && res[0] == 0 && res[1] == 0 && cap < (uint64_t)len * 1100 + 65536) continue;
        if (status != ZK_OK || res[1] != len || res[0] > cap || !c->shard.valid) return 1;
        rc = zk_inflate_shard_finish_device(c, d_window, d_out, d_result);
        if (rc != ZK_OK) return rc;
        ZK_CUDA(cudaMemcpyAsync(res, d_result, sizeof res, cudaMemcpyDeviceToHost, c->stream));
        ZK_CUDA(cudaStreamSynchronize(c->stream));
        if ((int)(int64_t)res[2] != ZK_STREAM_END) return 1;
        const size_t at = s->pending_out.size();
        s->pending_out.resize(at + res[0]);
        if (res[0]) {
            ZK_CUDA(cudaMemcpyAsync(s->pending_out.data() + at, d_out, res[0], cudaMemcpyDeviceToHost, c->stream));
            ZK_CUDA(cudaStreamSynchronize(c->stream));
        }
        c->d2h_bytes += (int64_t)res[0] + 64;
        if (s->container == 2) s->crc = res[0] ? zk_crc32(c, s->crc, s->pending_out.data() + at, res[0]) : s->crc;
        else s->adler = adler32_combine(s->adler, (uint32_t)res[3], res[0]);
        stream_note_output(s, s->pending_out.data() + at, res[0]);
        return 0;
    }
    return 1;
}

// The end of the stream: pending_in as raw deflate data that closes with a final block, the window as its dictionary.
// ZK_STREAM_END: decoded (*used = bytes of deflate data); ZK_BUF_ERROR: incomplete so far; anything else: an error.
int stream_decode_tail(zk_stream *s, uint64_t *used) {
    const size_t m = s->pending_in.size();
    uint64_t cap = (uint64_t)m * 4 + 65536;
    for (;;) {
        std::unique_ptr<uint8_t[]> buf(new uint8_t[cap]);  // (not a vector: no need to clear gigabytes for a try that may fail)
        uint64_t got = 0;
        uint32_t check = 1;
        const int rc = host_inflate(s->ctx, s->pending_in.data(), m, s->window.empty() ? nullptr : s->window.data(), (uint32_t)s->window.size(),
                                    buf.get(), cap, &got, used, 1, 0, &check, true, s->bit_off);
        if (rc == ZK_BUF_ERROR && got == 0 && *used == 0 && cap < (uint64_t)m * 1100 + 65536) {  // the output may not have fitted
            cap *= 16;
            continue;
        }
        if (rc != ZK_STREAM_END) return rc;
        if (s->container == 2) s->crc = got ? zk_crc32(s->ctx, s->crc, buf.get(), got) : s->crc;
        else s->adler = adler32_combine(s->adler, check, got);
        s->pending_out.insert(s->pending_out.end(), buf.get(), buf.get() + got);
        stream_note_output(s, buf.get(), got);
        return ZK_STREAM_END;
    }
}

// A stream without markers whose end has not arrived: decode the chunks between the entry points found in pending_in, all
// but the last (it may be cut), hand their bytes out and keep the rest -- which may begin inside a byte (bit_off).
// 0: progress; 1: none (too few entry points, more output than reserved, ...): the input goes on gathering.
int stream_decode_prefix(zk_stream *s) {
    const size_t m = s->pending_in.size();
    const uint64_t cap = (uint64_t)m * 5 + (1u << 20);
    std::unique_ptr<uint8_t[]> buf(new uint8_t[cap]);
    uint64_t got = 0, used_bits = 0;
    uint32_t check = 1;
    const int rc = host_inflate(s->ctx, s->pending_in.data(), m, s->window.empty() ? nullptr : s->window.data(), (uint32_t)s->window.size(),
                                buf.get(), cap, &got, &used_bits, 1, 0, &check, false, s->bit_off, true);
    if (rc != ZK_OK || got == 0 || (used_bits >> 3) > m) return 1;
    if (s->container == 2) s->crc = zk_crc32(s->ctx, s->crc, buf.get(), got);
    else s->adler = adler32_combine(s->adler, check, got);
    s->pending_out.insert(s->pending_out.end(), buf.get(), buf.get() + got);
    stream_note_output(s, buf.get(), got);
    s->pending_in.erase(s->pending_in.begin(), s->pending_in.begin() + (long)(used_bits >> 3));
    s->bit_off = (uint32_t)(used_bits & 7);
    s->plain_mode = true;  // from here on positions are bits: the marker logic is out
    s->scan_pos = s->pending_in.size();
    s->marker_end = 0;
    s->ctx->prefix_decodes++;
    return 0;
}

int stream_inflate_step(zk_stream *s, int flush, uint64_t *unread) {
    // ---- header ----
    if (!s->header_done) {
        const uint8_t *p = s->pending_in.data();
        const size_t m = s->pending_in.size();
        size_t body = 0;
        if (s->container == 0) {
            if (m < 2) return flush == ZK_FINISH ? ZK_BUF_ERROR : ZK_OK;
            const int cmf = p[0], flg = p[1];
            if ((cmf & 0x0F) != 8 || ((cmf >> 4) & 0x0F) > 7 || (((cmf << 8) | flg) % 31) != 0) return ZK_DATA_ERROR;  // IS:28-35
            body = 2;
            if (flg & 0x20) {  // FDICT, IS:36-43
                if (m < 6) return flush == ZK_FINISH ? ZK_BUF_ERROR : ZK_OK;
                if (!s->have_dict) return ZK_NEED_DICT;  // the input stays pending: zk_inflate_set_dictionary, then call again
                const uint32_t id = ((uint32_t)p[2] << 24) | ((uint32_t)p[3] << 16) | ((uint32_t)p[4] << 8) | p[5];
                const uint32_t mine = s->dict.empty() ? 1u : zk_adler32(s->ctx, 1u, s->dict.data(), s->dict.size());
                if (id != mine) return ZK_DATA_ERROR;
                const size_t keep = tmin<size_t>(s->dict.size(), 32768);
                s->window.assign(s->dict.end() - (long)keep, s->dict.end());
                body = 6;
            }
        } else if (s->container == 2) {
            const uint32_t h = gzip_header_len(p, m);
            if (h == 0) return ZK_DATA_ERROR;
            if (h == ~0u) return flush == ZK_FINISH ? ZK_BUF_ERROR : ZK_OK;
            body = h;
        } else if (s->have_dict) {  // raw deflate: zlib.h lets inflateSetDictionary prime the window
            const size_t keep = tmin<size_t>(s->dict.size(), 32768);
            s->window.assign(s->dict.end() - (long)keep, s->dict.end());
        }
        s->pending_in.erase(s->pending_in.begin(), s->pending_in.begin() + (long)body);
        s->header_done = true;
        s->adler = s->container == 2 ? 0u : 1u;
        s->crc = 0;
    }
    // ---- whole blocks up to the last sync marker that has arrived ----
    if (!s->body_done && !s->plain_mode) {
        const uint8_t *p = s->pending_in.data();
        const size_t m = s->pending_in.size();
        for (size_t i = s->scan_pos > 3 ? s->scan_pos - 3 : 0; i + 4 <= m;) {
            const void *f = memchr(p + i + 2, 0xFF, m - i - 2 - 1);
            if (!f) break;
            const size_t k = (size_t)(static_cast<const uint8_t *>(f) - p);  // p[k] == 0xFF, candidate marker p[k-2 .. k+1]
            if (p[k + 1] == 0xFF && p[k - 1] == 0 && p[k - 2] == 0) s->marker_end = k + 2;
            i = k - 1;
        }
        s->scan_pos = m;
        if (s->marker_end) {
            const int rc = stream_decode_marked(s, s->marker_end);
            if (rc < 0) return rc;
            if (rc == 0) {
                s->pending_in.erase(s->pending_in.begin(), s->pending_in.begin() + (long)s->marker_end);
                s->tail_attempt = 0;
                s->seen_marker = true;
            } else {
                s->plain_mode = true;  // not a block boundary (or damaged data): the end-of-stream decode below decides
            }
            s->scan_pos = s->pending_in.size();
            s->marker_end = 0;
        }
    }
    // ---- the end of the stream ----
    if (!s->body_done && !s->pending_in.empty() && flush != ZK_FINISH && !(s->seen_marker && !s->plain_mode)) {
        // no markers: a prefix up to the last entry point, whenever enough has gathered
        const size_t m = s->pending_in.size();
        const size_t least = s->ctx->opt_prefix_bytes ? (size_t)s->ctx->opt_prefix_bytes : (size_t)(32u << 20);
        if (m >= least && m >= 2 * s->prefix_attempt) {
            if (stream_decode_prefix(s) == 0) {
                s->prefix_attempt = 0;
                s->tail_attempt = s->pending_in.size();  // what is left is an unfinished chunk: try its end when it has grown
            } else {
                s->prefix_attempt = m;
            }
        }
    }
    if (!s->body_done && !s->pending_in.empty()) {
        const size_t m = s->pending_in.size();
        // In a stream of short marked blocks the rest behind the last marker is an incomplete block -- unless that block is
        // the final one (BFINAL is its first bit) or has outgrown any block of such a stream.  Without markers the end is
        // tried whenever the gathered input has doubled.
        const bool marked = s->seen_marker && !s->plain_mode;
        const bool looks_final = marked && (s->pending_in[0] & 1);
        if (flush == ZK_FINISH || looks_final || m >= tmax<size_t>(2 * s->tail_attempt, marked ? 65536 : 1)) {
            uint64_t used = 0;
            const int rc = stream_decode_tail(s, &used);
            if (rc == ZK_STREAM_END) {
                s->pending_in.erase(s->pending_in.begin(), s->pending_in.begin() + (long)used);
                s->body_done = true;
            } else if (rc == ZK_BUF_ERROR) {
                s->tail_attempt = m;
                if (flush == ZK_FINISH && s->pending_out.size() == s->out_pos) return ZK_BUF_ERROR;
            } else {
                return rc;
            }
        }
    } else if (!s->body_done && flush == ZK_FINISH && s->header_done && s->pending_out.size() == s->out_pos) {
        return ZK_BUF_ERROR;  // nothing left to decode and no final block seen
    }
    // ---- trailer (IS:414-424; RFC 1952: CRC-32 and ISIZE, little-endian) ----
    if (s->body_done && !s->finished) {
        const size_t need = s->container == 0 ? 4 : s->container == 2 ? 8 : 0;
        if (s->pending_in.size() < need) return flush == ZK_FINISH && s->pending_out.size() == s->out_pos ? ZK_BUF_ERROR : ZK_OK;
        const uint8_t *t = s->pending_in.data();
        bool good = true;
        if (s->container == 0) {
            good = ((((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3]) == s->adler);
        } else if (s->container == 2) {
            uint32_t crc = 0, isize = 0;
            for (int i = 0; i < 4; ++i) crc |= (uint32_t)t[i] << (8 * i), isize |= (uint32_t)t[4 + i] << (8 * i);
            good = crc == s->crc && isize == (uint32_t)s->decoded;
            s->adler = s->crc;
        }
        *unread = s->pending_in.size() - need;  // bytes behind the stream are not consumed
        s->pending_in.clear();
        s->finished = true;
        if (!good) return ZK_DATA_ERROR;
    }
    return ZK_OK;
}
}  // namespace

int zk_inflate(zk_stream *s, const uint8_t *in, uint64_t *in_len, uint8_t *out, uint64_t *out_len, int flush) {
    if (!s || s->mode != 1 || !in_len || !out_len) return ZK_STREAM_ERROR;
    const uint64_t avail_in = *in_len, avail_out = *out_len;
    *in_len = 0;
    *out_len = 0;
    if (s->sticky != ZK_OK) return s->sticky;
    try {
        if (!s->finished && avail_in) {
            if (!in) return ZK_STREAM_ERROR;
            s->pending_in.insert(s->pending_in.end(), in, in + avail_in);
            *in_len = avail_in;
            s->total_in += avail_in;
        }
        if (!s->finished) {
            uint64_t unread = 0;
            const int rc = stream_inflate_step(s, flush, &unread);
            if (unread) {
                s->total_in -= tmin<uint64_t>(s->total_in, unread);
                *in_len -= tmin<uint64_t>(*in_len, unread);
            }
            if (rc == ZK_NEED_DICT) return rc;
            if (rc != ZK_OK) {
                if (rc != ZK_BUF_ERROR) s->sticky = rc;
                // what was decoded before the error is still handed out below only for a failed checksum (IS:419-424)
                if (!(rc == ZK_DATA_ERROR && s->finished)) return rc;
            }
        }
    } catch (const std::bad_alloc &) {
        return ZK_MEM_ERROR;
    }
    const size_t have = s->pending_out.size() - s->out_pos;
    const size_t give = have < avail_out ? have : (size_t)avail_out;
    if (give) {
        if (!out) return ZK_STREAM_ERROR;
        memcpy(out, s->pending_out.data() + s->out_pos, give);
        s->out_pos += give;
        s->total_out += give;
        *out_len = give;
    }
    if (s->out_pos == s->pending_out.size()) {
        s->pending_out.clear();
        s->out_pos = 0;
    } else if (s->out_pos >= (8u << 20)) {  // drop what has been handed out now and then
        s->pending_out.erase(s->pending_out.begin(), s->pending_out.begin() + (long)s->out_pos);
        s->out_pos = 0;
    }
    if (s->sticky != ZK_OK) return s->sticky;
    if (s->finished && s->pending_out.empty()) return ZK_STREAM_END;
    // zlib.h: Z_FINISH on a stream that does not end in this input is Z_BUF_ERROR (once everything decodable is out)
    if (flush == ZK_FINISH && !s->finished && s->pending_out.empty()) return ZK_BUF_ERROR;
    if (!*in_len && !give) return ZK_BUF_ERROR;  // zlib.h: no progress possible
    return ZK_OK;
}

int zk_inflate_reset(zk_stream *s) {
    if (!s || s->mode != 1) return ZK_STREAM_ERROR;
    s->header_done = s->finished = false;
    s->total_in = s->total_out = 0;
    s->adler = 1;
    s->crc = 0;
    s->pending_in.clear();
    s->pending_out.clear();
    s->out_pos = 0;
    s->window.clear();
    s->scan_pos = s->marker_end = s->tail_attempt = s->prefix_attempt = 0;
    s->plain_mode = s->body_done = s->seen_marker = false;
    s->decoded = 0;
    s->bit_off = 0;
    s->sticky = ZK_OK;
    s->have_dict = false;  // zlib.h inflateReset: the stream starts over; a dictionary has to be supplied again
    s->dict.clear();
    return ZK_OK;
}

int zk_inflate_set_dictionary(zk_stream *s, const uint8_t *dict, uint32_t len) {
    // zlib.h inflateSetDictionary: called after zk_inflate returned ZK_NEED_DICT (or up front).  The reference itself has
    // no dictionary API and stops at Z_NEED_DICT (IS:36-43); this is the SURVEY.md 8(f) row built on the same priming
    // mechanism as the chunk halos.  A dictionary whose Adler-32 is not the stream's DICTID gives ZK_DATA_ERROR at decode.
    if (!s || s->mode != 1 || (!dict && len)) return ZK_STREAM_ERROR;
    try {
        s->dict.assign(dict, dict + len);
    } catch (const std::bad_alloc &) {
        return ZK_MEM_ERROR;
    }
    s->have_dict = true;
    return ZK_OK;
}

uint64_t zk_stream_total_in(const zk_stream *s) { return s ? s->total_in : 0; }
uint64_t zk_stream_total_out(const zk_stream *s) { return s ? s->total_out : 0; }
uint32_t zk_stream_adler(const zk_stream *s) { return s ? s->adler : 0; }

void zk_stream_end(zk_stream *s) { delete s; }

}  // extern "C"
#pragma GCC visibility pop
