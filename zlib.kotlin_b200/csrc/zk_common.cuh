// zk_common.cuh -- shared declarations of the sm_100a DEFLATE engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/zlibk_gpu.h"

namespace zk {

// ---- geometry -------------------------------------------------------------------------
constexpr int kWindow = 32768;     // DEFLATE window (DeflateStream.kt:573-574, InflateStream.kt:26)
constexpr int kMinMatch = 3;       // common/Constants.kt MIN_MATCH
constexpr int kMaxMatch = 258;     // common/Constants.kt MAX_MATCH
constexpr int kNumLitLen = 286;    // literal/length alphabet
constexpr int kNumDist = 30;       // distance alphabet
constexpr int kHistStride = 320;   // u32 counters per block: [0,286) lit/len, [288,318) dist
constexpr int kDistHistBase = 288;
constexpr int kGroup = 4096;       // LZ pipeline group (positions per pipeline step)
constexpr int kBlockBytesDefault = 32768;  // input span of one deflate block
constexpr int kMaxHeaderBytes = 320;       // dynamic block header: <= 17 + 57 + 316*(7+7) bits

// token / per-position match word: len << 16 | (dist - 1); len == 0 -> literal (low 8 bits = byte)
constexpr uint32_t kTokStart = 0x80000000u;   // parser mark: a token starts here
constexpr uint32_t kTokLiteral = 0x40000000u; // parser mark: emit the byte although a match exists

struct LevelParams {
    int chain;     // hash-chain steps per search
    int nice;      // stop searching once a match this long is found
    int lazy;      // one-step lazy evaluation
    int max_lazy;  // a match at least this long is taken without looking at the next position (zlib's max_lazy)
    int lazy2;     // > 0: a match shorter than this also gives way to a match at p + 2 that is at least two longer
    int far3;      // a length-3 match farther than this is dropped
    int far4;      // a length-4 match farther than this is dropped
    int min_len;   // 3, or 259 for Z_HUFFMAN_ONLY (no matches)
    int filtered;  // Z_FILTERED: matches of length <= 5 become literals
    int probe_adaptive;  // 1: probe short distances everywhere only while a sample of positions finds such matches
};

// Per-block record written by the Huffman kernel and consumed by the packer.
struct BlockPlan {
    uint32_t type;        // 0 stored, 1 fixed, 2 dynamic
    uint32_t header_bits; // bits of `header` (block header incl. BFINAL/BTYPE and code tables)
    uint32_t body_bits;   // token bits incl. end-of-block
    uint32_t out_bytes;   // bytes this block occupies in the stream (incl. the sync marker)
    uint32_t litlen_code[kNumLitLen]; // code | len << 16  (code already bit-reversed)
    uint32_t dist_code[kNumDist];
    uint8_t header[kMaxHeaderBytes];
};

template <typename T>
__host__ __device__ __forceinline__ T tmin(T a, T b) { return a < b ? a : b; }
template <typename T>
__host__ __device__ __forceinline__ T tmax(T a, T b) { return a > b ? a : b; }

// ---- error plumbing ---------------------------------------------------------------------
void set_last_error(const char *what, cudaError_t e);
#define ZK_CUDA(expr)                                      \
    do {                                                   \
        cudaError_t _e = (expr);                           \
        if (_e != cudaSuccess) {                           \
            ::zk::set_last_error(#expr, _e);               \
            return _e == cudaErrorMemoryAllocation ? ZK_MEM_ERROR : ZK_ERRNO; \
        }                                                  \
    } while (0)

// ---- device scratch arena ------------------------------------------------------------------
struct Arena {
    uint8_t *base = nullptr;
    size_t cap = 0, used = 0;
    int reserve(size_t bytes); // grow-only; invalidates earlier pointers (call before take)
    void reset() { used = 0; }
    void *take(size_t bytes, size_t align = 256);
    void release();
};

}  // namespace zk

struct zk_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;  // host-buffer entry points: H2D / D2H overlap with compute
    // deflate over several slabs: odd slabs run on alt_stream with their own scratch, so that the LZ kernel of the next
    // slab fills the SMs the tail of the current one leaves idle; slab_ev = {plan done (even, odd), fork, join}
    cudaStream_t alt_stream = nullptr;
    cudaEvent_t slab_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t opt_slab_overlap = 1;
    int64_t opt_slab_bytes = 0;                           // 0 = the built-in slab of a device-resident deflate
    int64_t opt_host_slab = 0, opt_host_first_slab = 0;  // 0 = the built-in sizes of the pipelined host call
    zk::Arena arena;       // device scratch (tokens, histograms, plans, staging)
    uint8_t *pinned = nullptr; // pinned host staging for small results
    size_t pinned_cap = 0;
    int64_t opt_chunk_bytes = 0; // 0 = auto
    int64_t opt_inflate_pipe = 0;     // 1: host-buffer inflate of a large stream with sync markers runs as a pipeline of slabs
                                      // (opt-in: measured slower than one pass, profiles/inflate_pipeline_r2.md)
    int64_t opt_auto_register = 0;    // 1: host-buffer calls page-lock caller buffers CUDA does not know yet and keep them locked
    void *registered[16] = {};        // ... those buffers (unlocked by zk_host_unregister or when the context goes)
    int num_registered = 0;
    bool stream_quick = false;        // set by the stream handle around a marked-body try: a body that runs out of input is answered at once
    int64_t opt_prefix_bytes = 0;     // stream handle: gathered bytes of a stream without markers from which a prefix is decoded (0 = 32 MiB)
    int64_t prefix_decodes = 0;       // counter: such prefix decodes that made progress
    int64_t opt_foreign_path = 1;     // 0: streams without sync markers are decoded by one warp (the round-1 behaviour)
    int64_t foreign_chunks = 0, foreign_merges = 0, quick_incomplete = 0;  // counters: chunks decoded through found entry points, false entry points merged
    int64_t opt_lane_pack = 0;        // 1: the lane decoder fills whole CTAs (few SMs per call) so that several calls run side by side
    int64_t opt_inflate_slab = 0;     // compressed bytes per slab of that pipeline, 0 = built-in
    zk_ctx *helpers[4] = {nullptr, nullptr, nullptr, nullptr};  // contexts of the pipeline's workers (own stream + arena)
    int64_t pipe_slabs = 0, pipe_fallbacks = 0;  // counters: slabs decoded by the pipeline, calls that fell back to one pass
    int64_t opt_block_bytes = zk::kBlockBytesDefault;
    int64_t opt_lz_chain = -1, opt_lz_max_lazy = -1, opt_lz_nice = -1;  // tuning overrides of the level table (-1 = off)
    int64_t launches = 0, h2d_bytes = 0, d2h_bytes = 0;
    // scratch of a sharded inflate between its two phases (pointers into the arena)
    struct ShardState {
        bool valid;
        void *d_chunks, *d_num, *d_verdict, *d_flag, *d_adler, *d_partials, *d_windows, *d_maps, *d_sym;
        uint32_t nch, group;
        uint64_t total;
        uint32_t dict_len;
        const uint8_t *d_in;
        uint64_t m;
    };
    ShardState shard = {false, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr, 0};
    // optional per-kernel-class device timing (option "timing"): event pairs recorded around launches
    int64_t opt_timing = 0;
    unsigned long long *lz_prof = nullptr;  // debug: per-unit cycle counters of lz77_kernel
    uint32_t *lz_dump_tokens = nullptr;     // debug: receives the token array of the (single) slab, u32[n]
    uint32_t *lz_dump_blocks = nullptr;     // debug: receives tok_begin[nb] then tok_count[nb]
    static constexpr int kTimedClasses = 6;  // 0 lz77, 1 huffman, 2 pack, 3 adler, 4 inflate-decode, 5 inflate-other
    struct TimedSpan {
        cudaEvent_t a, b;
        int cls;
    };
    TimedSpan *spans = nullptr;
    int span_count = 0, span_cap = 0;
    double class_ms[kTimedClasses] = {0, 0, 0, 0, 0, 0};
    int64_t class_launches[kTimedClasses] = {0, 0, 0, 0, 0, 0};
};

namespace zk {
// RAII event pair around one launch (no-op unless timing is on)
struct ScopedTimer {
    zk_ctx *c;
    int idx;
    ScopedTimer(zk_ctx *ctx, int cls);
    ~ScopedTimer();
};
void collect_timers(zk_ctx *c);  // after a stream sync: fold finished spans into class_ms
}  // namespace zk

namespace zk {

// ---- kernels' host launchers (each returns a ZK_* code) ------------------------------------
// K4: Adler-32 of d_buf[0,n) seeded with `seed`; result to d_out[0].  d_partials needs
// adler_partials_bytes(n) bytes of scratch.
size_t adler_partials_bytes(uint64_t n);
int launch_adler32(zk_ctx *c, uint32_t seed, const uint8_t *d_buf, uint64_t n, void *d_partials, uint32_t *d_out);

struct DeflateJob {
    const uint8_t *d_in;   // shard input (device)
    uint64_t n;
    const uint8_t *d_hist; // up to 32 KiB preceding the shard (device) or nullptr
    uint32_t hist_len;
    int level, strategy;
    uint8_t *d_out;        // body destination (device)
    uint64_t cap;
    bool framed;           // true: emit zlib header + final block + trailer around the body
    uint64_t *d_result;    // {out_len, adler, status, n}
    // optional pipelining hooks of the host-buffer entry point: the input is processed in slabs of `slab_bytes`;
    // before_slab(k) runs on the host before slab k's kernels are enqueued (wait for its H2D copy), after_slab(k,
    // d_cursor) after they are (d_cursor[0] = body bytes written so far, final for every slab up to k)
    uint64_t slab_bytes = 0;  // 0 = default
    uint64_t first_slab_bytes = 0;  // 0 = slab_bytes; a short first slab lets compute start before much has arrived
    void (*before_slab)(void *user, uint32_t k) = nullptr;
    void (*after_slab)(void *user, uint32_t k, const uint64_t *d_cursor) = nullptr;
    void *user = nullptr;
};
int run_deflate(zk_ctx *c, const DeflateJob &job);

struct InflateJob {
    const uint8_t *d_in;
    uint64_t m;
    uint8_t *d_out;
    uint64_t cap;
    uint64_t *d_result; // {out_len, consumed, status, adler}
    const uint8_t *d_dict = nullptr;  // preset dictionary (device) for a stream with FDICT, or nullptr
    uint32_t dict_len = 0;
    uint32_t dict_id = 0;             // Adler-32 of the whole dictionary: must equal the stream's DICTID
    int container = 0;                // 0 zlib, 1 raw deflate, 2 gzip member (header already parsed by the host), 3 shard body
    uint16_t *d_tail_map = nullptr;   // shard body: receives the shard's total tail map (u16[32768])
    uint32_t body_begin = 0;          // raw / gzip: offset of the deflate data in d_in
    // the streaming inflate tries the end of a stream while input is still arriving: when the chunks found by the entry-point
    // search are fine and only the last one runs out of input, answer ZK_BUF_ERROR at once instead of letting the one-warp
    // decoder walk the whole input to say the same
    bool quick_incomplete = false;
    // the stream handle, on a stream without markers whose end has not arrived: decode the chunks between found entry points
    // except the last (it may be cut), report the bit position where that one begins in result[1] and ZK_OK in result[2];
    // no progress possible -> {0, 0, ZK_BUF_ERROR}.  Raw container only.
    bool prefix = false;
    uint32_t begin_bit = 0;           // raw: the deflate data starts at this bit of byte body_begin (0..7)
};
int run_inflate(zk_ctx *c, const InflateJob &job);
int run_inflate_shard_finish(zk_ctx *c, const uint8_t *d_entry_window, uint8_t *d_out, uint64_t *d_result);
int launch_window_from_maps(zk_ctx *c, const uint16_t *d_maps, uint64_t stride_u16, uint32_t count, uint8_t *d_window);

// CRC-32 (gzip): same calling convention as launch_adler32
size_t crc_partials_bytes(uint64_t n);
int launch_crc32(zk_ctx *c, uint32_t seed, const uint8_t *d_buf, uint64_t n, void *d_partials, uint32_t *d_out);
uint32_t crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);

LevelParams level_params(int level, int strategy);
uint32_t adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2);

}  // namespace zk
