// ab_internal.h -- context object, error plumbing and launch helpers shared by the
// translation units of libastroburst_cuda.so.  Not part of the public ABI.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "astroburst_cuda.h"

constexpr int AB_NUM_COPY_STREAMS = 3;

struct ab_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;                        // default compute stream of the context
    cudaStream_t copy_streams[AB_NUM_COPY_STREAMS] = {};  // host-buffer pipelines
    cudaEvent_t copy_events[AB_NUM_COPY_STREAMS] = {};
    // device scratch reused by host-buffer entry points (grown on demand)
    void *d_scratch[AB_NUM_COPY_STREAMS] = {};
    size_t d_scratch_bytes[AB_NUM_COPY_STREAMS] = {};
    // small device + pinned host blocks for scalars (counters, stats partials)
    void *d_small = nullptr;
    void *h_small = nullptr;
    static constexpr size_t SMALL_BYTES = 2 << 20;
    int stack_variant = 0;     // AB_STACK_VARIANT: 0 TMA-staged (default), 2 plain coalesced loads
    int stack_mix = 0;         // AB_STACK_MIX: 0 default compare-exchange mix, 1..7 A/B variants (P = 32)
    bool stats_host_walks = false;  // AB_STATS_HOST_WALKS=1: percentile walks on the host (A/B; the sharded path always does)
    bool force_tma1d = false;  // AB_STACK_FORCE_TMA1D=1: per-frame bulk copies instead of the tensor map
    void *encode_tmap_fn = nullptr;  // cuTensorMapEncodeTiled, resolved lazily
    // work-queue counters of the stack kernels (one per launch in flight, round-robin)
    // Each launch takes slot i = seq % 256 of a ring of {work counter, finished-block ticket,
    // reject accumulator} triples that start at zero, and clears slot (i + 128) % 256 -- used 128
    // launches ago -- on its way: no memset launches around the stack kernel.
    static constexpr unsigned WORK_COUNTERS = 256;
    struct StackSlot {
        unsigned work, done;
        unsigned long long rejected;
    };
    StackSlot *d_work_counters = nullptr;
    std::atomic<unsigned> work_seq{0};
    // peer group (shard.cu): every rank's d_small mapped here, for in-kernel collectives over NVLink
    int peer_world = 0, peer_rank = 0;
    char *peer_small[AB_MAX_RANKS] = {};
    unsigned long long peer_epoch = 0;  // barriers issued so far (same sequence on every rank)
    // Calls on one context serialize on `mu`.  Recursive: host-buffer entry points hold it while
    // they call the _dev entry points, which take it themselves (they share d_small, the phase
    // correlation state and the scratch slots with every other call on the context).
    std::recursive_mutex mu;
    std::mutex err_mu;  // last_error is also written by argument checks that run before `mu`
    std::atomic<uint64_t> launches{0};
    std::string last_error;
};
using CtxLock = std::lock_guard<std::recursive_mutex>;

inline ab_status ab_fail(ab_ctx *ctx, ab_status st, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) {
        std::lock_guard<std::mutex> lk(ctx->err_mu);
        ctx->last_error = buf;
    }
    return st;
}

#define AB_CUDA(ctx, call)                                                                    \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return ab_fail((ctx), AB_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                \
                           cudaGetErrorString(e__), __FILE__, __LINE__);                      \
    } while (0)

#define AB_LAUNCH_CHECK(ctx)                                                                  \
    do {                                                                                      \
        (ctx)->launches.fetch_add(1, std::memory_order_relaxed);                              \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess)                                                               \
            return ab_fail((ctx), AB_ERR_CUDA, "kernel launch failed: %s (%s:%d)",            \
                           cudaGetErrorString(e__), __FILE__, __LINE__);                      \
    } while (0)

// RAII device guard: entry points run on the context's device without disturbing the caller's.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

ab_status ab_scratch_reserve(ab_ctx *ctx, int slot, size_t bytes);

// ---- stats.cu / stars.cu pieces used by preview.cu (no locking: the caller holds ctx->mu)
ab_status ab_scan_on(ab_ctx *ctx, const float *d, size_t len, cudaStream_t s, double *mn, double *mx,
                     double *sum, unsigned long long *count);
ab_status ab_detect_stars_on(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols,
                             double sigma_threshold, cudaStream_t s, std::vector<ab_star> &stars,
                             double *bg_median, double *bg_sigma);

// ---- phasecorr.cu
void ab_phasecorr_release(ab_ctx *ctx);
ab_status ab_phase_correlate_on(ab_ctx *ctx, const float *d_ref, const float *d_tgt, size_t rows,
                                size_t cols, cudaStream_t s, ab_phase_result *out);

// ---- shard.cu: all-reduce of regions of d_small over the peer group, enqueued on `s` (no host
// round trip, no NCCL): barrier, every rank sums the peers' copies in rank order, barrier.
// Offsets are bytes inside d_small; the result goes to `dst_off` (never in place).
constexpr size_t AB_SMALL_PEER_FLAGS = 896 << 10;   // u64 flag[AB_MAX_RANKS], then u32 error
constexpr size_t AB_SMALL_REDUCED = 1024 << 10;     // 512 KiB: reduced histograms A | B
struct PeerScalarOp {  // up to 8 scalars per call: kind 0 f64 sum, 1 u64 sum, 2 f64 max
    size_t src_off[8], dst_off[8];
    int kind[8];
    int n = 0;
    void add(size_t src, size_t dst, int k) {
        src_off[n] = src;
        dst_off[n] = dst;
        kind[n++] = k;
    }
};
// `count` u32 words (0 = none) and the scalars of `op` under ONE barrier pair
ab_status ab_peer_allreduce(ab_ctx *ctx, size_t src_off, size_t dst_off, size_t count,
                            const PeerScalarOp &op, cudaStream_t s);
ab_status ab_peer_check(ab_ctx *ctx, cudaStream_t s);  // after a synchronisation: did a barrier time out?

// ---- stack.cu
struct FrameTable {
    const float *p[AB_MAX_FRAMES];
};
// per-sample preprocessing / mode of the chunked deep kernel (stack_deep.cu)
struct DeepPre {
    const float *bias = nullptr;  // v -= bias[p]
    const float *dark = nullptr;  // v -= dark[p] * dark_ratio
    float dark_ratio = 1.0f;
    int median_only = 0;          // 1: upper median of the finite samples (calibration masters)
    int peer_frames = 0;          // 1: some frames live in a peer GPU's HBM (ab_stack_frame_sharded_dev)
};
ab_status ab_stack_launch_deep(ab_ctx *ctx, const FrameTable &ft, int n_frames, size_t npix,
                               float sigma_low, float sigma_high, unsigned max_iter,
                               const DeepPre &pre, float *d_out, unsigned long long *d_rejected,
                               cudaStream_t stream);
// fits.cu: decode_pixels (reader.rs:37-101) of a device-resident data unit, enqueued on `s`
ab_status ab_fits_decode_on(ab_ctx *ctx, const unsigned char *d, size_t n, int bitpix, double bscale,
                            double bzero, float *out, cudaStream_t s);
size_t ab_fits_bytes_per_pixel(int bitpix);  // 0 for an unsupported BITPIX
ab_status ab_stack_launch(ab_ctx *ctx, const FrameTable &ft, int n_frames, size_t npix,
                          float sigma_low, float sigma_high, unsigned max_iter, float *d_out,
                          unsigned long long *d_rejected, cudaStream_t stream, bool peer_frames = false);
