// stats.cu -- image statistics, 65,536-bin histograms and auto-STF parameters for sm_100a.
//
// Reference semantics (paths relative to src-tauri/src/):
//   core/imaging/stats.rs:11-13    is_valid_pixel (finite && > 1e-7)
//   core/imaging/stats.rs:15-41    compute_image_stats / _with_known_range (switch at > 4,000,000 px)
//   core/imaging/stats.rs:43-73    exact path: scan + exact_median_mut + exact_mad_mut
//   core/imaging/stats.rs:85-210   histogram path: value hist -> median refine + |v - coarse| hist
//                                  -> MAD refine, rank interpolation
//   core/imaging/stats.rs:378-444  build_histogram, downsample_histogram
//   core/imaging/stf.rs:13-47      auto_stf, mtf_balance (host scalar math)
//   math/median.rs:27-73           even-n averaging medians
//
// GPU mapping.  Every pass is one HBM sweep (float4 loads).  Bin indices use the reference's f64
// expression `((v as f64 - lo) * inv) as usize` evaluated on the FP64 pipe, so integer bin counts
// are bit-exact.  65,536-bin histograms: one 1,024-thread block per SM keeps the WHOLE histogram
// in shared memory as packed u16 counters (128 KiB), so every increment is a shared-memory atomic
// regardless of where the block's pixels fall; a counter reaching 0x8000 hands 0x8000 counts to
// the global u32 bin (L2-resident, 256 KiB).  The small rank/percentile walks over the
// histograms are host scalar code, as in the reference.  Exact medians (<= 4 M px) use an MSB-first
// radix select on the float bit patterns (valid pixels are positive, so bits are monotone).
// The f64 pixel sum is reduced block-partial by block-partial in a fixed order (deterministic;
// the reference's rayon reduction order is not).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#include <math_constants.h>

#include "ab_internal.h"

namespace {

constexpr float PADDING_THRESHOLD = 1e-7f;  // types/constants.rs:6
constexpr double MAD_TO_SIGMA = 1.4826;     // types/constants.rs:7
constexpr unsigned HIST_BINS = 65536;       // stats.rs:8
constexpr int ST_THREADS = 512;
constexpr int HT_THREADS = 1024;            // histogram passes: one block per SM
constexpr unsigned HALF_FLUSH = 0x8000u;    // packed u16 counter -> global u32 bin hand-over
// One increment of packed half `a` (shared-window address of the table in `pbase`): a
// shared-space atomic on a precomputed 32-bit address -- the generic-pointer form re-derives the
// shared window base (S2R + 3 integer instructions) for every pixel.
__device__ __forceinline__ void packed_inc(unsigned pbase, unsigned a, unsigned *__restrict__ histA) {
    const unsigned addr = pbase + ((a & ~1u) << 1);
    const unsigned odd = a & 1u;
    const unsigned inc = odd ? 0x10000u : 1u;
    unsigned old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(inc) : "memory");
    const unsigned half = odd ? (old >> 16) : (old & 0xffffu);
    if (half == HALF_FLUSH - 1u) {
        atomicAdd(&histA[a], HALF_FLUSH);
        const unsigned dec = 0u - (odd ? (HALF_FLUSH << 16) : HALF_FLUSH);
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(dec) : "memory");
    }
}

__device__ __forceinline__ bool valid_px(float v) {
    return (fabsf(v) < CUDART_INF_F) && v > PADDING_THRESHOLD;
}

struct ScanPartial {
    double sum;
    unsigned long long count;
    float mn, mx;
};

// ---- pass: min / max / sum / count over valid pixels (stats.rs:212-257)
__global__ void __launch_bounds__(ST_THREADS)
scan_kernel(const float *__restrict__ data, size_t len, ScanPartial *__restrict__ partials) {
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    double sum = 0.0;
    unsigned long long cnt = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(data) & 15) == 0;
    const size_t n4 = vec_ok ? len / 4 : 0;
    auto acc = [&](float v) {
        if (valid_px(v)) {
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
            sum += (double)v;
            ++cnt;
        }
    };
    auto four = [&](const float4 &v) { acc(v.x); acc(v.y); acc(v.z); acc(v.w); };
    const float4 *d4 = reinterpret_cast<const float4 *>(data);
    size_t i = tid;
    for (; i + 3 * stride < n4; i += 4 * stride) {  // four 16-byte loads in flight per thread
        const float4 v0 = __ldcs(d4 + i);
        const float4 v1 = __ldcs(d4 + i + stride);
        const float4 v2 = __ldcs(d4 + i + 2 * stride);
        const float4 v3 = __ldcs(d4 + i + 3 * stride);
        four(v0); four(v1); four(v2); four(v3);
    }
    for (; i < n4; i += stride) four(__ldcs(d4 + i));
    for (size_t j = n4 * 4 + tid; j < len; j += stride) acc(data[j]);

    __shared__ ScanPartial sh[ST_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        sum += __shfl_down_sync(0xffffffffu, sum, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = {sum, cnt, mn, mx};
    __syncthreads();
    if (threadIdx.x == 0) {
        ScanPartial r = sh[0];
        for (int w = 1; w < ST_THREADS / 32; ++w) {
            r.sum += sh[w].sum;
            r.count += sh[w].count;
            r.mn = fminf(r.mn, sh[w].mn);
            r.mx = fmaxf(r.mx, sh[w].mx);
        }
        partials[blockIdx.x] = r;
    }
}

// ---- generic 65,536-bin histogram pass over the valid pixels.  The classifier F says at compile
// time what a pass collects: kSum (f64 sum + count), kA (bin_a -> histogram A, shared memory),
// kB (bin_b -> sparse refine histogram B, global atomics), kBelow (is_below -> counter).
struct HistPartial {
    double sum;
    unsigned long long count;
    unsigned long long below;
};

template <typename F>
__global__ void __launch_bounds__(HT_THREADS, 1)
hist_kernel(const float *__restrict__ data, size_t len, F f, unsigned *__restrict__ histA,
            unsigned *__restrict__ histB, HistPartial *__restrict__ partials) {
    // histogram A lives whole in shared memory as 65,536 packed u16 counters (128 KiB, one block
    // per SM): every increment is a shared-memory atomic wherever the block's pixels fall.  A
    // counter that reaches 0x8000 moves 0x8000 counts to the global u32 bin; at most
    // HT_THREADS - 1 further increments can land before the subtraction, so a half never wraps.
    extern __shared__ unsigned packed[];
    for (unsigned i = threadIdx.x; i < HIST_BINS / 2; i += HT_THREADS) packed[i] = 0;
    __syncthreads();
    const unsigned pbase = (unsigned)__cvta_generic_to_shared(packed);

    const bool vec_ok = (reinterpret_cast<uintptr_t>(data) & 15) == 0;
    const size_t n4 = vec_ok ? len / 4 : 0;
    const size_t stride = (size_t)gridDim.x * HT_THREADS;
    const size_t tid = (size_t)blockIdx.x * HT_THREADS + threadIdx.x;
    double sum = 0.0;
    unsigned cnt = 0, below = 0;  // per thread < 2^32 (len < 2^32)

    // one early-out for invalid pixels, then straight-line code selected at compile time by the
    // classifier's flags (kSum / kA / kB / kBelow): no per-pixel flag tests
    auto one = [&](float v) {
        if (!valid_px(v)) return;
        const double dv = (double)v;
        if constexpr (F::kSum) {
            sum += dv;
            ++cnt;
        }
        if constexpr (F::kA) {
            const unsigned a = f.bin_a(v, dv);
            packed_inc(pbase, a, histA);
        }
        if constexpr (F::kB) {
            unsigned b;
            if (f.bin_b(v, dv, b)) atomicAdd(&histB[b], 1u);
        }
        if constexpr (F::kBelow) below += f.is_below(v) ? 1u : 0u;
    };
    auto four = [&](const float4 &v) { one(v.x); one(v.y); one(v.z); one(v.w); };
    const float4 *d4 = reinterpret_cast<const float4 *>(data);
    size_t i = tid;
    for (; i + 3 * stride < n4; i += 4 * stride) {  // four 16-byte loads in flight per thread
        const float4 v0 = __ldcs(d4 + i);
        const float4 v1 = __ldcs(d4 + i + stride);
        const float4 v2 = __ldcs(d4 + i + 2 * stride);
        const float4 v3 = __ldcs(d4 + i + 3 * stride);
        four(v0); four(v1); four(v2); four(v3);
    }
    for (; i < n4; i += stride) four(__ldcs(d4 + i));
    for (size_t j = n4 * 4 + tid; j < len; j += stride) one(data[j]);
    __syncthreads();
    for (unsigned w = threadIdx.x; w < HIST_BINS / 2; w += HT_THREADS) {
        const unsigned c = packed[w];
        if (c & 0xffffu) atomicAdd(&histA[2 * w], c & 0xffffu);
        if (c >> 16) atomicAdd(&histA[2 * w + 1], c >> 16);
    }
    // deterministic per-block partial of the f64 sum
    __shared__ HistPartial sh[HT_THREADS / 32];
    unsigned long long cnt64 = cnt, below64 = below;
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, o);
        cnt64 += __shfl_down_sync(0xffffffffu, cnt64, o);
        below64 += __shfl_down_sync(0xffffffffu, below64, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = {sum, cnt64, below64};
    __syncthreads();
    if (threadIdx.x == 0) {
        HistPartial r = sh[0];
        for (int w = 1; w < HT_THREADS / 32; ++w) {
            r.sum += sh[w].sum;
            r.count += sh[w].count;
            r.below += sh[w].below;
        }
        partials[blockIdx.x] = r;
    }
}

// Same pass with the classifier read from device memory: the device-chained statistics path
// (stats_hist_chained) computes each pass's parameters in a one-block "walk" kernel on the same
// stream, so the sweeps follow each other without a host round trip.
template <typename F>
__global__ void __launch_bounds__(HT_THREADS, 1)
hist_ctl_kernel(const float *__restrict__ data, size_t len, const F *__restrict__ f_dev,
                const int *__restrict__ skip, unsigned *__restrict__ histA,
                unsigned *__restrict__ histB, HistPartial *__restrict__ partials);

// `((x) as usize).min(last)` for x = f64.  The f64 -> u32 conversion saturates like Rust's `as`
// (negative and NaN -> 0, >= 2^32 -> u32::MAX), so the clamp to the last bin is exact for every x.
__device__ __forceinline__ unsigned bin_of(double x) {
    return min(__double2uint_rz(x), HIST_BINS - 1);
}

struct ValueHistF {  // stats.rs:259-297 scan_sum_and_hist; also build_histogram :393-400
    static constexpr bool kSum = true, kA = true, kB = false, kBelow = false;
    double lo, inv;
    __device__ unsigned bin_a(float, double dv) const { return bin_of((dv - lo) * inv); }
    __device__ bool bin_b(float, double, unsigned &) const { return false; }
    __device__ bool is_below(float) const { return false; }
};
struct RefineDevF {  // stats.rs:119-146
    static constexpr bool kSum = false, kA = true, kB = true, kBelow = false;
    double bin_lo, bin_hi, refine_inv, dev_inv;
    float coarse_med;
    __device__ unsigned bin_a(float v, double) const {
        const float d = fabsf(v - coarse_med);  // f32 subtraction
        return bin_of((double)d * dev_inv);
    }
    __device__ bool bin_b(float, double vf, unsigned &b) const {
        if (!(vf >= bin_lo && vf < bin_hi)) return false;
        b = bin_of((vf - bin_lo) * refine_inv);
        return true;
    }
    __device__ bool is_below(float) const { return false; }
};
struct MadRefineF {  // stats.rs:164-193
    static constexpr bool kSum = false, kA = false, kB = true, kBelow = true;
    double region_lo, refine_inv;
    float med, lo_f32, hi_f32;
    __device__ unsigned bin_a(float, double) const { return 0; }
    __device__ bool bin_b(float v, double, unsigned &b) const {
        const float dev = fabsf(v - med);
        if (!(dev >= lo_f32 && dev < hi_f32)) return false;
        b = bin_of(((double)dev - region_lo) * refine_inv);
        return true;
    }
    __device__ bool is_below(float v) const { return fabsf(v - med) < lo_f32; }
};

template <typename F>
__global__ void __launch_bounds__(HT_THREADS, 1)
hist_ctl_kernel(const float *__restrict__ data, size_t len, const F *__restrict__ f_dev,
                const int *__restrict__ skip, unsigned *__restrict__ histA,
                unsigned *__restrict__ histB, HistPartial *__restrict__ partials) {
    extern __shared__ unsigned packed[];
    if (*skip) {  // no valid pixel: nothing to classify
        if (threadIdx.x == 0) partials[blockIdx.x] = {0.0, 0ull, 0ull};
        return;
    }
    const F f = *f_dev;
    for (unsigned i = threadIdx.x; i < HIST_BINS / 2; i += HT_THREADS) packed[i] = 0;
    __syncthreads();
    const unsigned pbase = (unsigned)__cvta_generic_to_shared(packed);
    const bool vec_ok = (reinterpret_cast<uintptr_t>(data) & 15) == 0;
    const size_t n4 = vec_ok ? len / 4 : 0;
    const size_t stride = (size_t)gridDim.x * HT_THREADS;
    const size_t tid = (size_t)blockIdx.x * HT_THREADS + threadIdx.x;
    double sum = 0.0;
    unsigned cnt = 0, below = 0;
    auto one = [&](float v) {
        if (!valid_px(v)) return;
        const double dv = (double)v;
        if constexpr (F::kSum) {
            sum += dv;
            ++cnt;
        }
        if constexpr (F::kA) {
            const unsigned a = f.bin_a(v, dv);
            packed_inc(pbase, a, histA);
        }
        if constexpr (F::kB) {
            unsigned b;
            if (f.bin_b(v, dv, b)) atomicAdd(&histB[b], 1u);
        }
        if constexpr (F::kBelow) below += f.is_below(v) ? 1u : 0u;
    };
    auto four = [&](const float4 &v) { one(v.x); one(v.y); one(v.z); one(v.w); };
    const float4 *d4 = reinterpret_cast<const float4 *>(data);
    size_t i = tid;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        const float4 v0 = __ldcs(d4 + i);
        const float4 v1 = __ldcs(d4 + i + stride);
        const float4 v2 = __ldcs(d4 + i + 2 * stride);
        const float4 v3 = __ldcs(d4 + i + 3 * stride);
        four(v0); four(v1); four(v2); four(v3);
    }
    for (; i < n4; i += stride) four(__ldcs(d4 + i));
    for (size_t j = n4 * 4 + tid; j < len; j += stride) one(data[j]);
    __syncthreads();
    if constexpr (F::kA) {
        for (unsigned w = threadIdx.x; w < HIST_BINS / 2; w += HT_THREADS) {
            const unsigned c = packed[w];
            if (c & 0xffffu) atomicAdd(&histA[2 * w], c & 0xffffu);
            if (c >> 16) atomicAdd(&histA[2 * w + 1], c >> 16);
        }
    }
    __shared__ HistPartial sh[HT_THREADS / 32];
    unsigned long long cnt64 = cnt, below64 = below;
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, o);
        cnt64 += __shfl_down_sync(0xffffffffu, cnt64, o);
        below64 += __shfl_down_sync(0xffffffffu, below64, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = {sum, cnt64, below64};
    __syncthreads();
    if (threadIdx.x == 0) {
        HistPartial r = sh[0];
        for (int w = 1; w < HT_THREADS / 32; ++w) {
            r.sum += sh[w].sum;
            r.count += sh[w].count;
            r.below += sh[w].below;
        }
        partials[blockIdx.x] = r;
    }
}

// ---- device-side percentile walks (stats.rs:299-353 on the device, one 1,024-thread block).
// All of it is integer counting plus a handful of IEEE f64 operations (no contraction:
// -fmad=false), so the values equal the host walk's bit for bit.
struct StatsCtl {
    ValueHistF f1;
    RefineDevF f2;
    MadRefineF f3;
    double global_min, global_max, range, bin_width, mean;
    double median_bin_lo, refine_range, median, dev_bw, mad_region_lo, mad_refine_range;
    unsigned long long total_valid, half_count, count_before_median;
    // per-pass scalars, folded from the block partials and (sharded images) all-reduced in place
    double mm[2];                   // {-min, max} of the valid pixels: one MAX all-reduce covers both
    unsigned long long scan_count;  // valid pixels seen by the scan pass
    double hsum;                    // pass 1: f64 sum of the valid pixels
    unsigned long long hcnt[2];     // {valid count, count below the MAD region}
    // the same after the cross-rank reduction (what the walks read; = the local values on one GPU)
    double mm_r[2];
    unsigned long long scan_count_r;
    double hsum_r;
    unsigned long long hcnt_r[2];
    int empty;
    ab_image_stats out;
};
// cross-rank reduction of DEVICE buffers on the stream (sharded images); null = single GPU.
// dtype 0 f64 / 1 u64 / 2 u32, op 0 sum / 1 min / 2 max (ab_allreduce_dev_fn in the header)
struct DevReducer {
    ab_allreduce_dev_fn fn = nullptr;
    void *user = nullptr;
    bool peer = false;  // reduce over the peer group (shard.cu) instead of the callback
    ab_status run(ab_ctx *ctx, void *d_buf, size_t count, int dtype, int op, cudaStream_t s) const {
        if (!fn) return AB_OK;
        if (fn(user, d_buf, count, dtype, op, (void *)s) != 0)
            return ab_fail(ctx, AB_ERR_INVALID, "device all-reduce callback failed");
        return AB_OK;
    }
};
// one warp: lane l folds partials l, l + 32, ... in order, then a fixed xor tree -- deterministic,
// and a few microseconds instead of a serial walk over up to 444 global-memory entries
__global__ void fold_scan_kernel(const ScanPartial *__restrict__ sp, unsigned n, StatsCtl *ctl) {
    double lo = DBL_MAX, hi = -DBL_MAX;
    unsigned long long c = 0;
    for (unsigned i = threadIdx.x; i < n; i += 32) {
        if (sp[i].count) {
            lo = fmin(lo, (double)sp[i].mn);
            hi = fmax(hi, (double)sp[i].mx);
        }
        c += sp[i].count;
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (threadIdx.x == 0) {
        ctl->mm[0] = ctl->mm_r[0] = -lo;
        ctl->mm[1] = ctl->mm_r[1] = hi;
        ctl->scan_count = ctl->scan_count_r = c;
    }
}
__global__ void fold_hist_kernel(const HistPartial *__restrict__ hp, unsigned n, StatsCtl *ctl) {
    double s = 0.0;
    unsigned long long c = 0, b = 0;
    for (unsigned i = threadIdx.x; i < n; i += 32) {
        s += hp[i].sum;
        c += hp[i].count;
        b += hp[i].below;
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (threadIdx.x == 0) {
        ctl->hsum = ctl->hsum_r = s;
        ctl->hcnt[0] = ctl->hcnt_r[0] = c;
        ctl->hcnt[1] = ctl->hcnt_r[1] = b;
    }
}
constexpr int WK_THREADS = 1024, WK_PER = HIST_BINS / WK_THREADS;

// first bin i with cum(i) >= target (cum inclusive): *idx, *cum = cum(i), *cnt = hist[i].
// All threads of the block call it; returns false when the total stays below the target.
__device__ bool block_find_cum(const unsigned *__restrict__ hist, unsigned long long target,
                               unsigned *idx, unsigned long long *cum, unsigned long long *cnt) {
    __shared__ unsigned long long part[WK_THREADS];
    __shared__ unsigned long long res[3];
    __shared__ int found;
    const unsigned b0 = threadIdx.x * WK_PER;
    unsigned long long sum = 0;
    for (int j = 0; j < WK_PER; ++j) sum += hist[b0 + j];
    part[threadIdx.x] = sum;
    if (threadIdx.x == 0) found = 0;
    __syncthreads();
    {  // exclusive scan of the 1,024 chunk sums: warp scans + a scan of the 32 warp totals
        __shared__ unsigned long long wtot[32];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        unsigned long long inc = sum;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = wtot[lane], winc = w;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long v = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += v;
            }
            wtot[lane] = winc - w;
        }
        __syncthreads();
        part[threadIdx.x] = wtot[warp] + inc - sum;
    }
    __syncthreads();
    const unsigned long long before = part[threadIdx.x];
    if (before < target && before + sum >= target) {  // exactly one chunk crosses the target
        unsigned long long run = before;
        for (int j = 0; j < WK_PER; ++j) {
            const unsigned long long c = hist[b0 + j];
            run += c;
            if (run >= target) {
                res[0] = b0 + j;
                res[1] = run;
                res[2] = c;
                found = 1;
                break;
            }
        }
    }
    __syncthreads();
    const bool ok = found != 0;
    if (ok) {
        *idx = (unsigned)res[0];
        *cum = res[1];
        *cnt = res[2];
    }
    __syncthreads();
    return ok;
}
__device__ __forceinline__ double walk_frac(unsigned long long cum, unsigned long long target,
                                            unsigned long long count) {
    const unsigned long long overshoot = cum - target;
    return count > 0 ? 1.0 - ((double)overshoot / (double)count) : 0.5;
}
// resolve_rank_in_hist, stats.rs:333-353 (all threads)
__device__ double block_resolve_rank(const unsigned *hist, unsigned long long rank, double region_lo,
                                     double sub_bin_width) {
    if (rank == 0) return region_lo;
    unsigned idx = 0;
    unsigned long long cum = 0, cnt = 0;
    if (block_find_cum(hist, rank, &idx, &cum, &cnt))
        return region_lo + ((double)idx + walk_frac(cum, rank, cnt)) * sub_bin_width;
    return region_lo + (double)HIST_BINS * sub_bin_width;
}

// after the scan pass (or with a caller-known range): range, bin width, pass-1 classifier
__global__ void walk_range_kernel(int known, double kmin, double kmax, StatsCtl *ctl) {
    double lo = kmin, hi = kmax;
    int empty = 0;
    if (!known) {
        lo = -ctl->mm_r[0];
        hi = ctl->mm_r[1];
        empty = ctl->scan_count_r == 0;
    }
    ctl->empty = empty;
    ctl->global_min = lo;
    ctl->global_max = hi;
    const double range = fmax(hi - lo, 1e-30);
    ctl->range = range;
    ctl->bin_width = range / (double)HIST_BINS;
    ctl->f1.lo = lo;
    ctl->f1.inv = (double)HIST_BINS / range;
}
// after pass 1 (value histogram + sum): median bin, coarse median, pass-2 classifier
__global__ void __launch_bounds__(WK_THREADS)
walk_value_kernel(const unsigned *__restrict__ histA, StatsCtl *ctl) {
    const unsigned long long total = ctl->hcnt_r[0];
    const double gsum = ctl->hsum_r;
    const int was_empty = ctl->empty;
    __syncthreads();  // every thread has read `empty` before thread 0 may overwrite it
    if (was_empty || total == 0) {
        if (threadIdx.x == 0) ctl->empty = 1;
        return;
    }
    const double bin_width = ctl->bin_width, gmin = ctl->global_min, range = ctl->range;
    const unsigned long long target = (unsigned long long)ceil((double)total * 0.5);
    unsigned idx = HIST_BINS - 1;
    unsigned long long cum = 0, cnt = 0;
    const bool ok = block_find_cum(histA, target, &idx, &cum, &cnt);
    if (threadIdx.x == 0) {
        // find_percentile_bin (:299-309) / interpolate_percentile (:311-331)
        const unsigned median_bin = ok ? idx : HIST_BINS - 1;
        unsigned long long before = 0;
        if (ok) before = cum - cnt;
        else
            for (unsigned i = 0; i < HIST_BINS - 1; ++i) before += histA[i];
        const double coarse = ok ? gmin + ((double)idx + walk_frac(cum, target, cnt)) * bin_width
                                 : gmin + (double)HIST_BINS * bin_width;
        const double lo = gmin + (double)median_bin * bin_width, hi = lo + bin_width;
        const double refine_range = fmax(hi - lo, 1e-30);
        ctl->total_valid = total;
        ctl->mean = gsum / (double)total;
        ctl->half_count = target;
        ctl->count_before_median = before;
        ctl->median_bin_lo = lo;
        ctl->refine_range = refine_range;
        ctl->dev_bw = range / (double)HIST_BINS;
        ctl->f2.bin_lo = lo;
        ctl->f2.bin_hi = hi;
        ctl->f2.refine_inv = (double)HIST_BINS / refine_range;
        ctl->f2.dev_inv = (double)HIST_BINS / range;
        ctl->f2.coarse_med = (float)coarse;
    }
}
// after pass 2 (deviation histogram A + median-refine histogram B): median, MAD region, pass-3 classifier
__global__ void __launch_bounds__(WK_THREADS)
walk_refine_kernel(const unsigned *__restrict__ histA, const unsigned *__restrict__ histB, StatsCtl *ctl) {
    if (ctl->empty) return;
    const unsigned long long half = ctl->half_count, before = ctl->count_before_median;
    const unsigned long long rank = half > before ? half - before : 0;
    const double median = block_resolve_rank(histB, rank, ctl->median_bin_lo,
                                             ctl->refine_range / (double)HIST_BINS);
    unsigned idx = HIST_BINS - 1;
    unsigned long long cum = 0, cnt = 0;
    const bool ok = block_find_cum(histA, half, &idx, &cum, &cnt);  // ceil(total * 0.5) again
    if (threadIdx.x == 0) {
        const unsigned mad_bin = ok ? idx : HIST_BINS - 1;
        const unsigned expand_lo = mad_bin > 0 ? mad_bin - 1 : 0;
        const unsigned expand_hi = min(mad_bin + 2, HIST_BINS);
        const double dev_bw = ctl->dev_bw;
        const double region_lo = (double)expand_lo * dev_bw, region_hi = (double)expand_hi * dev_bw;
        const double mrr = fmax(region_hi - region_lo, 1e-30);
        ctl->median = median;
        ctl->mad_region_lo = region_lo;
        ctl->mad_refine_range = mrr;
        ctl->f3.region_lo = region_lo;
        ctl->f3.refine_inv = (double)HIST_BINS / mrr;
        ctl->f3.med = (float)median;
        ctl->f3.lo_f32 = (float)region_lo;
        ctl->f3.hi_f32 = (float)region_hi;
    }
}
// after pass 3 (MAD-refine histogram B + count below): MAD, sigma, the ImageStats
__global__ void __launch_bounds__(WK_THREADS)
walk_mad_kernel(const unsigned *__restrict__ histB, StatsCtl *ctl) {
    if (ctl->empty) {
        if (threadIdx.x == 0) ctl->out = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0};  // ImageStats::default(), like stats_default
        return;
    }
    const unsigned long long half = ctl->half_count, below = ctl->hcnt_r[1];
    const unsigned long long rank = half > below ? half - below : 0;
    const double mad = block_resolve_rank(histB, rank, ctl->mad_region_lo,
                                          ctl->mad_refine_range / (double)HIST_BINS);
    if (threadIdx.x == 0) {
        ab_image_stats o;
        o.min = ctl->global_min;
        o.max = ctl->global_max;
        o.median = ctl->median;
        o.mad = mad;
        o.sigma = fmax(mad * MAD_TO_SIGMA, 1e-30);
        o.mean = ctl->mean;
        o.valid_count = ctl->total_valid;
        ctl->out = o;
    }
}

// ---- radix select on float bit patterns (exact path)
// key(v) = bits(v) for mode 0, bits(|v - center|) for mode 1; only valid pixels take part.
__global__ void __launch_bounds__(ST_THREADS)
radix_hist_kernel(const float *__restrict__ data, size_t len, int mode, float center,
                  unsigned prefix, unsigned prefix_mask, int shift, unsigned *__restrict__ hist256) {
    __shared__ unsigned sh[256];
    if (threadIdx.x < 256) sh[threadIdx.x] = 0;
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const float v = data[i];
        if (!valid_px(v)) continue;
        const unsigned key = __float_as_uint(mode ? fabsf(v - center) : v);
        if ((key & prefix_mask) == prefix) atomicAdd(&sh[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 256 && sh[threadIdx.x]) atomicAdd(&hist256[threadIdx.x], sh[threadIdx.x]);
}

unsigned grid_for(ab_ctx *ctx, size_t len, int per_thread) {
    size_t blocks = (len / per_thread + ST_THREADS - 1) / ST_THREADS + 1;
    const size_t cap = (size_t)ctx->sm_count * 3;
    return (unsigned)(blocks > cap ? cap : blocks);
}

// k-th smallest (0-based) key among the valid pixels; 4 passes of 8 bits.
ab_status radix_select(ab_ctx *ctx, const float *d, size_t len, int mode, float center,
                       unsigned long long k, cudaStream_t s, float *out) {
    unsigned *d_hist = reinterpret_cast<unsigned *>(static_cast<char *>(ctx->d_small) + 65536);
    unsigned *h_hist = reinterpret_cast<unsigned *>(static_cast<char *>(ctx->h_small) + 65536);
    unsigned prefix = 0, mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        AB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, 256 * sizeof(unsigned), s));
        radix_hist_kernel<<<grid_for(ctx, len, 8), ST_THREADS, 0, s>>>(d, len, mode, center, prefix,
                                                                      mask, shift, d_hist);
        AB_LAUNCH_CHECK(ctx);
        AB_CUDA(ctx, cudaMemcpyAsync(h_hist, d_hist, 256 * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
        AB_CUDA(ctx, cudaStreamSynchronize(s));
        unsigned long long cum = 0;
        int b = 0;
        for (; b < 256; ++b) {
            if (cum + h_hist[b] > k) break;
            cum += h_hist[b];
        }
        if (b == 256) return ab_fail(ctx, AB_ERR_CUDA, "radix select: rank out of range");
        k -= cum;
        prefix |= (unsigned)b << shift;
        mask |= 0xffu << shift;
    }
    float r;
    std::memcpy(&r, &prefix, 4);
    *out = r;
    return AB_OK;
}

// exact_median_mut / median_f32_mut semantics on the selected order statistics (median.rs:27-63)
ab_status select_median_pair(ab_ctx *ctx, const float *d, size_t len, int mode, float center,
                             unsigned long long n, cudaStream_t s, float *left, float *right) {
    const unsigned long long mid = n / 2;
    ab_status st = radix_select(ctx, d, len, mode, center, mid, s, right);
    if (st != AB_OK) return st;
    *left = *right;
    if (n % 2 == 0) st = radix_select(ctx, d, len, mode, center, mid - 1, s, left);
    return st;
}

void stats_default(ab_image_stats *o) { std::memset(o, 0, sizeof(*o)); }

// stats.rs:299-353 on host copies of the histograms
size_t find_percentile_bin(const unsigned *hist, size_t bins, unsigned long long total, double pct) {
    const unsigned long long target = (unsigned long long)std::ceil((double)total * pct);
    unsigned long long cum = 0;
    for (size_t i = 0; i < bins; ++i) {
        cum += hist[i];
        if (cum >= target) return i;
    }
    return bins - 1;
}
double interpolate_percentile(const unsigned *hist, size_t bins, unsigned long long total,
                              double pct, double data_min, double bin_width) {
    const unsigned long long target = (unsigned long long)std::ceil((double)total * pct);
    unsigned long long cum = 0;
    for (size_t i = 0; i < bins; ++i) {
        const unsigned long long count = hist[i];
        cum += count;
        if (cum >= target) {
            const unsigned long long overshoot = cum - target;
            const double frac = count > 0 ? 1.0 - ((double)overshoot / (double)count) : 0.5;
            return data_min + ((double)i + frac) * bin_width;
        }
    }
    return data_min + (double)bins * bin_width;
}
double resolve_rank_in_hist(const unsigned *hist, size_t bins, unsigned long long rank,
                            double region_lo, double sub_bin_width) {
    if (rank == 0) return region_lo;
    unsigned long long cum = 0;
    for (size_t i = 0; i < bins; ++i) {
        const unsigned long long count = hist[i];
        cum += count;
        if (cum >= rank) {
            const unsigned long long overshoot = cum - rank;
            const double frac = count > 0 ? 1.0 - ((double)overshoot / (double)count) : 0.5;
            return region_lo + ((double)i + frac) * sub_bin_width;
        }
    }
    return region_lo + (double)bins * sub_bin_width;
}

// device scratch layout inside ctx->d_small / h_small (1 MiB each):
//   [0, 64 KiB)        partials (<= 444 blocks x 24 B)
//   [64 KiB, 68 KiB)   radix 256-bin histogram
//   [128 KiB, 384 KiB) histogram A (65,536 x u32)
//   [384 KiB, 640 KiB) histogram B
struct Small {
    char *d, *h;
    unsigned *dA, *dB, *hA, *hB;
    explicit Small(ab_ctx *ctx)
        : d(static_cast<char *>(ctx->d_small)), h(static_cast<char *>(ctx->h_small)) {
        dA = reinterpret_cast<unsigned *>(d + (128 << 10));
        dB = reinterpret_cast<unsigned *>(d + (384 << 10));
        hA = reinterpret_cast<unsigned *>(h + (128 << 10));
        hB = reinterpret_cast<unsigned *>(h + (384 << 10));
    }
};

template <typename F>
ab_status run_hist(ab_ctx *ctx, const float *d, size_t len, F f, bool wantA, bool wantB,
                   cudaStream_t s, double *sum, unsigned long long *count,
                   unsigned long long *below) {
    Small sm(ctx);
    size_t blocks = (len / 16 + HT_THREADS - 1) / HT_THREADS + 1;
    if (blocks > (size_t)ctx->sm_count) blocks = (size_t)ctx->sm_count;
    const unsigned grid = (unsigned)blocks;
    constexpr int smem = (int)(HIST_BINS / 2 * sizeof(unsigned));
    AB_CUDA(ctx, cudaMemsetAsync(sm.dA, 0, 2 * HIST_BINS * sizeof(unsigned), s));
    AB_CUDA(ctx, cudaFuncSetAttribute(hist_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    hist_kernel<F><<<grid, HT_THREADS, smem, s>>>(
        d, len, f, sm.dA, sm.dB, reinterpret_cast<HistPartial *>(sm.d));
    AB_LAUNCH_CHECK(ctx);
    AB_CUDA(ctx, cudaMemcpyAsync(sm.h, sm.d, grid * sizeof(HistPartial), cudaMemcpyDeviceToHost, s));
    if (wantA)
        AB_CUDA(ctx, cudaMemcpyAsync(sm.hA, sm.dA, HIST_BINS * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    if (wantB)
        AB_CUDA(ctx, cudaMemcpyAsync(sm.hB, sm.dB, HIST_BINS * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    AB_CUDA(ctx, cudaStreamSynchronize(s));
    const HistPartial *hp = reinterpret_cast<const HistPartial *>(sm.h);
    double sm_sum = 0.0;
    unsigned long long c = 0, b = 0;
    for (unsigned i = 0; i < grid; ++i) {
        sm_sum += hp[i].sum;
        c += hp[i].count;
        b += hp[i].below;
    }
    if (sum) *sum = sm_sum;
    if (count) *count = c;
    if (below) *below = b;
    return AB_OK;
}

ab_status run_scan(ab_ctx *ctx, const float *d, size_t len, cudaStream_t s, double *mn, double *mx,
                   double *sum, unsigned long long *count) {
    Small sm(ctx);
    const unsigned grid = grid_for(ctx, len, 16);
    scan_kernel<<<grid, ST_THREADS, 0, s>>>(d, len, reinterpret_cast<ScanPartial *>(sm.d));
    AB_LAUNCH_CHECK(ctx);
    AB_CUDA(ctx, cudaMemcpyAsync(sm.h, sm.d, grid * sizeof(ScanPartial), cudaMemcpyDeviceToHost, s));
    AB_CUDA(ctx, cudaStreamSynchronize(s));
    const ScanPartial *sp = reinterpret_cast<const ScanPartial *>(sm.h);
    double lo = DBL_MAX, hi = -DBL_MAX, tot = 0.0;
    unsigned long long c = 0;
    for (unsigned i = 0; i < grid; ++i) {
        if (sp[i].count) {
            lo = std::fmin(lo, (double)sp[i].mn);
            hi = std::fmax(hi, (double)sp[i].mx);
        }
        tot += sp[i].sum;
        c += sp[i].count;
    }
    *mn = lo;
    *mx = hi;
    *sum = tot;
    *count = c;
    return AB_OK;
}

// Cross-rank reduction hook for row-band sharded images: every histogram / counter
// the algorithm consumes is summed over ranks before the (identical, replicated) host walk.
// dtype: 0 f64 sum, 1 u64 sum, 2 f64 min, 3 f64 max.  Null = single GPU.
struct Reducer {
    ab_allreduce_fn fn = nullptr;
    void *user = nullptr;
    bool on() const { return fn != nullptr; }
    ab_status run(ab_ctx *ctx, void *buf, size_t count, int dtype) const {
        if (!fn) return AB_OK;
        if (fn(user, buf, count, dtype) != 0)
            return ab_fail(ctx, AB_ERR_INVALID, "all-reduce callback failed");
        return AB_OK;
    }
};

// sums a u32 host histogram over ranks (through u64, like the reference's u64 bins)
ab_status reduce_hist(ab_ctx *ctx, const Reducer &rd, unsigned *h, std::vector<unsigned long long> &wide) {
    wide.resize(HIST_BINS);
    for (size_t i = 0; i < HIST_BINS; ++i) wide[i] = h[i];
    return rd.run(ctx, wide.data(), HIST_BINS, 1);
}

size_t find_percentile_bin64(const unsigned long long *hist, size_t bins, unsigned long long total,
                             double pct) {
    const unsigned long long target = (unsigned long long)std::ceil((double)total * pct);
    unsigned long long cum = 0;
    for (size_t i = 0; i < bins; ++i) {
        cum += hist[i];
        if (cum >= target) return i;
    }
    return bins - 1;
}
double interpolate_percentile64(const unsigned long long *hist, size_t bins, unsigned long long total,
                                double pct, double data_min, double bin_width) {
    const unsigned long long target = (unsigned long long)std::ceil((double)total * pct);
    unsigned long long cum = 0;
    for (size_t i = 0; i < bins; ++i) {
        const unsigned long long count = hist[i];
        cum += count;
        if (cum >= target) {
            const unsigned long long overshoot = cum - target;
            const double frac = count > 0 ? 1.0 - ((double)overshoot / (double)count) : 0.5;
            return data_min + ((double)i + frac) * bin_width;
        }
    }
    return data_min + (double)bins * bin_width;
}
double resolve_rank_in_hist64(const unsigned long long *hist, size_t bins, unsigned long long rank,
                              double region_lo, double sub_bin_width) {
    if (rank == 0) return region_lo;
    unsigned long long cum = 0;
    for (size_t i = 0; i < bins; ++i) {
        const unsigned long long count = hist[i];
        cum += count;
        if (cum >= rank) {
            const unsigned long long overshoot = cum - rank;
            const double frac = count > 0 ? 1.0 - ((double)overshoot / (double)count) : 0.5;
            return region_lo + ((double)i + frac) * sub_bin_width;
        }
    }
    return region_lo + (double)bins * sub_bin_width;
}

// stats.rs:85-210
ab_status hist_core(ab_ctx *ctx, const float *d, size_t len, double global_min, double global_max,
                    cudaStream_t s, ab_image_stats *out, const Reducer &rd = Reducer()) {
    Small sm(ctx);
    std::vector<unsigned long long> hA, hB;
    const double range = std::fmax(global_max - global_min, 1e-30);
    const double bin_width = range / (double)HIST_BINS;
    const double inv_bin = (double)HIST_BINS / range;

    double global_sum = 0.0;
    unsigned long long total_valid = 0;
    ab_status st = run_hist(ctx, d, len, ValueHistF{global_min, inv_bin}, true, false, s,
                            &global_sum, &total_valid, nullptr);
    if (st != AB_OK) return st;
    if ((st = reduce_hist(ctx, rd, sm.hA, hA)) != AB_OK) return st;
    if ((st = rd.run(ctx, &global_sum, 1, 0)) != AB_OK) return st;
    if ((st = rd.run(ctx, &total_valid, 1, 1)) != AB_OK) return st;
    if (total_valid == 0) {
        stats_default(out);
        return AB_OK;
    }
    const unsigned long long n = total_valid;
    const double mean = global_sum / (double)n;
    const unsigned long long half_count = (unsigned long long)std::ceil((double)total_valid * 0.5);

    const size_t median_bin = find_percentile_bin64(hA.data(), HIST_BINS, total_valid, 0.5);
    unsigned long long count_before_median = 0;
    for (size_t i = 0; i < median_bin; ++i) count_before_median += hA[i];
    const double median_bin_lo = global_min + (double)median_bin * bin_width;
    const double median_bin_hi = median_bin_lo + bin_width;
    const double coarse_median =
        interpolate_percentile64(hA.data(), HIST_BINS, total_valid, 0.5, global_min, bin_width);

    const double dev_range = range;
    const double dev_bw = dev_range / (double)HIST_BINS;
    const double dev_inv = (double)HIST_BINS / dev_range;
    const float coarse_med_f32 = (float)coarse_median;
    const double refine_range = std::fmax(median_bin_hi - median_bin_lo, 1e-30);
    const double refine_inv = (double)HIST_BINS / refine_range;

    st = run_hist(ctx, d, len,
                  RefineDevF{median_bin_lo, median_bin_hi, refine_inv, dev_inv, coarse_med_f32},
                  true, true, s, nullptr, nullptr, nullptr);
    if (st != AB_OK) return st;
    if ((st = reduce_hist(ctx, rd, sm.hA, hA)) != AB_OK) return st;
    if ((st = reduce_hist(ctx, rd, sm.hB, hB)) != AB_OK) return st;
    const unsigned long long median_rank_in_bin =
        half_count > count_before_median ? half_count - count_before_median : 0;
    const double median_refine_bw = refine_range / (double)HIST_BINS;
    const double median = resolve_rank_in_hist64(hB.data(), HIST_BINS, median_rank_in_bin,
                                                 median_bin_lo, median_refine_bw);

    const size_t mad_bin = find_percentile_bin64(hA.data(), HIST_BINS, total_valid, 0.5);
    const size_t expand_lo = mad_bin > 0 ? mad_bin - 1 : 0;
    const size_t expand_hi = std::min(mad_bin + 2, (size_t)HIST_BINS);
    const double mad_region_lo = (double)expand_lo * dev_bw;
    const double mad_region_hi = (double)expand_hi * dev_bw;

    const float exact_med_f32 = (float)median;
    const double mad_refine_range = std::fmax(mad_region_hi - mad_region_lo, 1e-30);
    const double mad_refine_inv = (double)HIST_BINS / mad_refine_range;
    unsigned long long count_below = 0;
    st = run_hist(ctx, d, len,
                  MadRefineF{mad_region_lo, mad_refine_inv, exact_med_f32, (float)mad_region_lo,
                             (float)mad_region_hi},
                  false, true, s, nullptr, nullptr, &count_below);
    if (st != AB_OK) return st;
    if ((st = reduce_hist(ctx, rd, sm.hB, hB)) != AB_OK) return st;
    if ((st = rd.run(ctx, &count_below, 1, 1)) != AB_OK) return st;
    const unsigned long long mad_rank = half_count > count_below ? half_count - count_below : 0;
    const double mad_refine_bw = mad_refine_range / (double)HIST_BINS;
    const double mad =
        resolve_rank_in_hist64(hB.data(), HIST_BINS, mad_rank, mad_region_lo, mad_refine_bw);

    out->min = global_min;
    out->max = global_max;
    out->mean = mean;
    out->sigma = std::fmax(mad * MAD_TO_SIGMA, 1e-30);
    out->median = median;
    out->mad = mad;
    out->valid_count = n;
    return AB_OK;
}

// stats.rs:75-210 with every percentile walk on the device: scan, walk, 3 x (sweep, walk) queued
// back to back on `s`, ONE synchronisation for the 56-byte result.  `known`: the caller's range
// (compute_image_stats_with_known_range, :25-41) replaces the scan pass.  With a DevReducer the
// image is a row band of a larger one: the folded scalars and the histograms are all-reduced in
// place on the stream between a sweep and its walk, every rank then walks identical data.
ab_status stats_hist_chained(ab_ctx *ctx, const float *d, size_t len, bool known, double kmin,
                             double kmax, cudaStream_t s, ab_image_stats *out,
                             const DevReducer &rd = DevReducer()) {
    Small sm(ctx);
    // control block + host mirror of the result in the small scratch blocks ([640 KiB, +4 KiB))
    constexpr size_t ctl_off = 640 << 10, offA = 128 << 10, offB = 384 << 10;
    StatsCtl *ctl = reinterpret_cast<StatsCtl *>(sm.d + ctl_off);
    ab_image_stats *h_out = reinterpret_cast<ab_image_stats *>(sm.h + ctl_off);
    // the histograms the walks read: the local ones, or (peer group) the reduced copies
    unsigned *wA = rd.peer ? reinterpret_cast<unsigned *>(sm.d + AB_SMALL_REDUCED) : sm.dA;
    unsigned *wB = rd.peer ? wA + HIST_BINS : sm.dB;
    auto hist_scalars = [&]() {
        PeerScalarOp op;
        op.add(ctl_off + offsetof(StatsCtl, hsum), ctl_off + offsetof(StatsCtl, hsum_r), 0);
        op.add(ctl_off + offsetof(StatsCtl, hcnt), ctl_off + offsetof(StatsCtl, hcnt_r), 1);
        op.add(ctl_off + offsetof(StatsCtl, hcnt) + 8, ctl_off + offsetof(StatsCtl, hcnt_r) + 8, 1);
        return op;
    };
    ScanPartial *d_sp = reinterpret_cast<ScanPartial *>(sm.d);
    HistPartial *d_hp = reinterpret_cast<HistPartial *>(sm.d + (32 << 10));
    ab_status st;
    if (!known) {
        const unsigned n_sp = len ? grid_for(ctx, len, 16) : 0;
        if (n_sp) {
            scan_kernel<<<n_sp, ST_THREADS, 0, s>>>(d, len, d_sp);
            AB_LAUNCH_CHECK(ctx);
        }
        fold_scan_kernel<<<1, 32, 0, s>>>(d_sp, n_sp, ctl);
        AB_LAUNCH_CHECK(ctx);
        if (rd.peer) {
            PeerScalarOp op;
            op.add(ctl_off + offsetof(StatsCtl, mm), ctl_off + offsetof(StatsCtl, mm_r), 2);
            op.add(ctl_off + offsetof(StatsCtl, mm) + 8, ctl_off + offsetof(StatsCtl, mm_r) + 8, 2);
            op.add(ctl_off + offsetof(StatsCtl, scan_count), ctl_off + offsetof(StatsCtl, scan_count_r), 1);
            if ((st = ab_peer_allreduce(ctx, 0, 0, 0, op, s)) != AB_OK) return st;
        }
        if ((st = rd.run(ctx, ctl->mm_r, 2, 0, 2, s)) != AB_OK) return st;
        if ((st = rd.run(ctx, &ctl->scan_count_r, 1, 1, 0, s)) != AB_OK) return st;
    }
    walk_range_kernel<<<1, 1, 0, s>>>(known ? 1 : 0, kmin, kmax, ctl);
    AB_LAUNCH_CHECK(ctx);
    size_t blocks = (len / 16 + HT_THREADS - 1) / HT_THREADS + 1;
    if (blocks > (size_t)ctx->sm_count) blocks = (size_t)ctx->sm_count;
    const unsigned grid = len ? (unsigned)blocks : 0;
    constexpr int smem = (int)(HIST_BINS / 2 * sizeof(unsigned));
    AB_CUDA(ctx, cudaFuncSetAttribute(hist_ctl_kernel<ValueHistF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    AB_CUDA(ctx, cudaFuncSetAttribute(hist_ctl_kernel<RefineDevF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    AB_CUDA(ctx, cudaFuncSetAttribute(hist_ctl_kernel<MadRefineF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const size_t hist_bytes = 2 * HIST_BINS * sizeof(unsigned);
    // ---- pass 1: value histogram + sum
    AB_CUDA(ctx, cudaMemsetAsync(sm.dA, 0, hist_bytes, s));
    if (grid) {
        hist_ctl_kernel<ValueHistF><<<grid, HT_THREADS, smem, s>>>(d, len, &ctl->f1, &ctl->empty, sm.dA, sm.dB, d_hp);
        AB_LAUNCH_CHECK(ctx);
    }
    fold_hist_kernel<<<1, 32, 0, s>>>(d_hp, grid, ctl);
    AB_LAUNCH_CHECK(ctx);
    if (rd.peer && (st = ab_peer_allreduce(ctx, offA, AB_SMALL_REDUCED, HIST_BINS, hist_scalars(), s)) != AB_OK)
        return st;
    if ((st = rd.run(ctx, sm.dA, HIST_BINS, 2, 0, s)) != AB_OK) return st;
    if ((st = rd.run(ctx, &ctl->hsum_r, 1, 0, 0, s)) != AB_OK) return st;
    if ((st = rd.run(ctx, ctl->hcnt_r, 2, 1, 0, s)) != AB_OK) return st;
    walk_value_kernel<<<1, WK_THREADS, 0, s>>>(wA, ctl);
    AB_LAUNCH_CHECK(ctx);
    // ---- pass 2: deviation histogram (A) + median-refine histogram (B)
    AB_CUDA(ctx, cudaMemsetAsync(sm.dA, 0, hist_bytes, s));
    if (grid) {
        hist_ctl_kernel<RefineDevF><<<grid, HT_THREADS, smem, s>>>(d, len, &ctl->f2, &ctl->empty, sm.dA, sm.dB, d_hp);
        AB_LAUNCH_CHECK(ctx);
    }
    if (rd.peer && (st = ab_peer_allreduce(ctx, offA, AB_SMALL_REDUCED, 2 * HIST_BINS, PeerScalarOp(), s)) != AB_OK)
        return st;
    if ((st = rd.run(ctx, sm.dA, 2 * HIST_BINS, 2, 0, s)) != AB_OK) return st;  // A and B are adjacent
    walk_refine_kernel<<<1, WK_THREADS, 0, s>>>(wA, wB, ctl);
    AB_LAUNCH_CHECK(ctx);
    // ---- pass 3: MAD-refine histogram (B) + count below the region
    AB_CUDA(ctx, cudaMemsetAsync(sm.dB, 0, HIST_BINS * sizeof(unsigned), s));
    if (grid) {
        hist_ctl_kernel<MadRefineF><<<grid, HT_THREADS, smem, s>>>(d, len, &ctl->f3, &ctl->empty, sm.dA, sm.dB, d_hp);
        AB_LAUNCH_CHECK(ctx);
    }
    fold_hist_kernel<<<1, 32, 0, s>>>(d_hp, grid, ctl);
    AB_LAUNCH_CHECK(ctx);
    if (rd.peer && (st = ab_peer_allreduce(ctx, offB, AB_SMALL_REDUCED + HIST_BINS * sizeof(unsigned), HIST_BINS,
                                           hist_scalars(), s)) != AB_OK)
        return st;
    if ((st = rd.run(ctx, sm.dB, HIST_BINS, 2, 0, s)) != AB_OK) return st;
    if ((st = rd.run(ctx, ctl->hcnt_r, 2, 1, 0, s)) != AB_OK) return st;
    walk_mad_kernel<<<1, WK_THREADS, 0, s>>>(wB, ctl);
    AB_LAUNCH_CHECK(ctx);
    AB_CUDA(ctx, cudaMemcpyAsync(h_out, &ctl->out, sizeof(ab_image_stats), cudaMemcpyDeviceToHost, s));
    AB_CUDA(ctx, cudaStreamSynchronize(s));
    *out = *h_out;
    return rd.peer ? ab_peer_check(ctx, s) : AB_OK;
}

// stats.rs:43-73
ab_status stats_exact(ab_ctx *ctx, const float *d, size_t len, cudaStream_t s, ab_image_stats *out) {
    double mn, mx, sum;
    unsigned long long n;
    ab_status st = run_scan(ctx, d, len, s, &mn, &mx, &sum, &n);
    if (st != AB_OK) return st;
    if (n == 0) {
        stats_default(out);
        return AB_OK;
    }
    float left, right;
    st = select_median_pair(ctx, d, len, 0, 0.0f, n, s, &left, &right);
    if (st != AB_OK) return st;
    const double median = (n % 2 == 0) ? ((double)left + (double)right) / 2.0 : (double)right;
    st = select_median_pair(ctx, d, len, 1, (float)median, n, s, &left, &right);
    if (st != AB_OK) return st;
    const float mad_f32 = (n % 2 == 0) ? (left + right) / 2.0f : right;
    const double mad = (double)mad_f32;
    out->min = mn;
    out->max = mx;
    out->mean = sum / (double)n;
    out->sigma = std::fmax(mad * MAD_TO_SIGMA, 1e-30);
    out->median = median;
    out->mad = mad;
    out->valid_count = n;
    return AB_OK;
}

ab_status stats_hist(ab_ctx *ctx, const float *d, size_t len, cudaStream_t s, ab_image_stats *out) {
    double mn, mx, sum;
    unsigned long long n;
    if (!ctx->stats_host_walks) return stats_hist_chained(ctx, d, len, false, 0.0, 0.0, s, out);
    ab_status st = run_scan(ctx, d, len, s, &mn, &mx, &sum, &n);  // scan_minmax, stats.rs:212-231
    if (st != AB_OK) return st;
    if (n == 0) {
        stats_default(out);
        return AB_OK;
    }
    return hist_core(ctx, d, len, mn, mx, s, out);
}

// mode: 0 compute_image_stats, 1 exact, 2 histogram, 3 with_known_range
ab_status stats_dispatch(ab_ctx *ctx, const float *d, size_t len, int mode, double kmin, double kmax,
                         cudaStream_t s, ab_image_stats *out) {
    if (len == 0) {
        stats_default(out);
        return AB_OK;
    }
    if (len >= ((size_t)1 << 32))
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "images of 2^32 pixels or more are not supported");
    switch (mode) {
        case 1:
            return stats_exact(ctx, d, len, s, out);
        case 2:
            return stats_hist(ctx, d, len, s, out);
        case 3:
            if (len <= 4000000) return stats_exact(ctx, d, len, s, out);
            if (!std::isfinite(kmin) || !std::isfinite(kmax) || kmin >= kmax)
                return stats_hist(ctx, d, len, s, out);
            if (!ctx->stats_host_walks) return stats_hist_chained(ctx, d, len, true, kmin, kmax, s, out);
            return hist_core(ctx, d, len, kmin, kmax, s, out);
        default:
            return len > 4000000 ? stats_hist(ctx, d, len, s, out) : stats_exact(ctx, d, len, s, out);
    }
}

ab_status upload_plane(ab_ctx *ctx, const float *h, size_t len, float **d) {
    ab_status st = ab_scratch_reserve(ctx, 0, len * 4 + 256);
    if (st != AB_OK) return st;
    *d = static_cast<float *>(ctx->d_scratch[0]);
    AB_CUDA(ctx, cudaMemcpyAsync(*d, h, len * 4, cudaMemcpyHostToDevice, ctx->stream));
    return AB_OK;
}

}  // namespace

// min / max / sum / count over the valid pixels (stats.rs:212-257); min = DBL_MAX when count == 0
ab_status ab_scan_on(ab_ctx *ctx, const float *d, size_t len, cudaStream_t s, double *mn, double *mx,
                     double *sum, unsigned long long *count) {
    return run_scan(ctx, d, len, s, mn, mx, sum, count);
}

extern "C" {

ab_status ab_image_stats_dev(ab_ctx *ctx, const float *d_data, size_t len, int mode,
                             double known_min, double known_max, ab_image_stats *out,
                             void *stream) {
    if (!ctx || !out || (len && !d_data)) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    return stats_dispatch(ctx, d_data, len, mode, known_min, known_max, (cudaStream_t)stream, out);
}

ab_status ab_compute_image_stats(ab_ctx *ctx, const float *data, size_t len, int mode,
                                 double known_min, double known_max, ab_image_stats *out) {
    if (!ctx || !out || (len && !data)) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    if (len == 0) {
        stats_default(out);
        return AB_OK;
    }
    float *d = nullptr;
    ab_status st = upload_plane(ctx, data, len, &d);
    if (st != AB_OK) return st;
    return stats_dispatch(ctx, d, len, mode, known_min, known_max, ctx->stream, out);
}

// Row-band sharded compute_image_stats (histogram algorithm, stats.rs:75-210): `d_data` is this
// rank's band, the callback sums / mins / maxes host buffers over all ranks.  Every rank returns
// the same ImageStats.  Bit-identical to the single-GPU result except `mean` (f64 sum order, G9).
ab_status ab_image_stats_sharded_dev(ab_ctx *ctx, const float *d_data, size_t len,
                                     ab_allreduce_fn allreduce, void *user, int use_known_range,
                                     double known_min, double known_max, ab_image_stats *out,
                                     void *stream) {
    if (!ctx || !out || !allreduce || (len && !d_data)) return AB_ERR_INVALID;
    if (len >= ((size_t)1 << 32))
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "bands of 2^32 pixels or more are not supported");
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    Reducer rd;
    rd.fn = allreduce;
    rd.user = user;
    double mn = known_min, mx = known_max;
    const bool known = use_known_range && std::isfinite(known_min) && std::isfinite(known_max) &&
                       known_min < known_max;
    if (!known) {
        double sum = 0.0;
        unsigned long long n = 0;
        mn = DBL_MAX;
        mx = -DBL_MAX;
        if (len) {
            ab_status st = run_scan(ctx, d_data, len, s, &mn, &mx, &sum, &n);
            if (st != AB_OK) return st;
        }
        ab_status st = rd.run(ctx, &mn, 1, 2);
        if (st != AB_OK) return st;
        if ((st = rd.run(ctx, &mx, 1, 3)) != AB_OK) return st;
        if ((st = rd.run(ctx, &n, 1, 1)) != AB_OK) return st;
        if (n == 0) {
            stats_default(out);
            return AB_OK;
        }
    }
    return hist_core(ctx, d_data, len, mn, mx, s, out, rd);
}

// The same with every reduction on the device and on the stream (NCCL through the callback): no
// host round trip between the sweeps, one synchronisation for the result.
ab_status ab_image_stats_sharded_chained_dev(ab_ctx *ctx, const float *d_data, size_t len,
                                             ab_allreduce_dev_fn allreduce, void *user,
                                             int use_known_range, double known_min, double known_max,
                                             ab_image_stats *out, void *stream) {
    if (!ctx || !out || !allreduce || (len && !d_data)) return AB_ERR_INVALID;
    if (len >= ((size_t)1 << 32))
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "bands of 2^32 pixels or more are not supported");
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    DevReducer rd;
    rd.fn = allreduce;
    rd.user = user;
    const bool known = use_known_range && std::isfinite(known_min) && std::isfinite(known_max) &&
                       known_min < known_max;
    return stats_hist_chained(ctx, d_data, len, known, known_min, known_max, (cudaStream_t)stream, out, rd);
}

// The same over the peer group: no callback, the reductions are kernels reading the peers' blocks.
ab_status ab_image_stats_peer_dev(ab_ctx *ctx, const float *d_data, size_t len, int use_known_range,
                                  double known_min, double known_max, ab_image_stats *out, void *stream) {
    if (!ctx || !out || (len && !d_data)) return AB_ERR_INVALID;
    if (len >= ((size_t)1 << 32))
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "bands of 2^32 pixels or more are not supported");
    CtxLock lk(ctx->mu);
    if (ctx->peer_world < 1) return ab_fail(ctx, AB_ERR_INVALID, "no peer group (ab_peer_connect)");
    DeviceGuard g(ctx->device);
    DevReducer rd;
    rd.peer = true;
    const bool known = use_known_range && std::isfinite(known_min) && std::isfinite(known_max) &&
                       known_min < known_max;
    return stats_hist_chained(ctx, d_data, len, known, known_min, known_max, (cudaStream_t)stream, out, rd);
}

// build_histogram, stats.rs:378-421: bins = n_bins u32, edges = n_bins + 1 f64 (host outputs)
static ab_status histogram_impl(ab_ctx *ctx, const float *d, size_t len, size_t n_bins, double dmin,
                                double dmax, uint32_t *bins, double *edges, cudaStream_t s) {
    if (n_bins != HIST_BINS)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "only 65536-bin histograms (HISTOGRAM_BINS) are built");
    const double range = dmax - dmin;
    if (range < 1e-10) {  // stats.rs:379-387
        std::memset(bins, 0, n_bins * sizeof(uint32_t));
        for (size_t i = 0; i <= n_bins; ++i) edges[i] = dmin;
        return AB_OK;
    }
    const double inv = (double)n_bins / range;
    Small sm(ctx);
    ab_status st = run_hist(ctx, d, len, ValueHistF{dmin, inv}, true, false, s, nullptr, nullptr, nullptr);
    if (st != AB_OK) return st;
    std::memcpy(bins, sm.hA, n_bins * sizeof(uint32_t));
    const double step = range / (double)n_bins;
    for (size_t i = 0; i <= n_bins; ++i) edges[i] = dmin + (double)i * step;
    return AB_OK;
}
ab_status ab_histogram_u32_dev(ab_ctx *ctx, const float *d_data, size_t len, size_t n_bins,
                               double dmin, double dmax, uint32_t *bins, double *edges,
                               void *stream) {
    if (!ctx || !bins || !edges || (len && !d_data)) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    return histogram_impl(ctx, d_data, len, n_bins, dmin, dmax, bins, edges, (cudaStream_t)stream);
}
ab_status ab_histogram_u32(ab_ctx *ctx, const float *data, size_t len, size_t n_bins, double dmin,
                           double dmax, uint32_t *bins, double *edges) {
    if (!ctx || !bins || !edges || (len && !data)) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    float *d = nullptr;
    if (len) {
        ab_status st = upload_plane(ctx, data, len, &d);
        if (st != AB_OK) return st;
    }
    return histogram_impl(ctx, d, len, n_bins, dmin, dmax, bins, edges, ctx->stream);
}

// downsample_histogram, stats.rs:423-444 (host scalar loop over 65,536 bins)
ab_status ab_downsample_histogram(const uint32_t *src, size_t src_len, size_t target_bins,
                                  uint32_t *out) {
    if (!src || !out) return AB_ERR_INVALID;
    if (target_bins >= src_len) {
        std::memcpy(out, src, src_len * sizeof(uint32_t));
        return AB_OK;
    }
    const double ratio = (double)src_len / (double)target_bins;
    for (size_t i = 0; i < target_bins; ++i) {
        const size_t start = (size_t)((double)i * ratio);
        const size_t end = std::min((size_t)((double)(i + 1) * ratio), src_len);
        uint32_t sum = 0;
        for (size_t j = start; j < end; ++j) {
            const uint64_t t = (uint64_t)sum + src[j];
            sum = t > 0xffffffffull ? 0xffffffffu : (uint32_t)t;
        }
        out[i] = sum;
    }
    return AB_OK;
}

// auto_stf + mtf_balance, stf.rs:13-47 (host scalar math)
ab_status ab_auto_stf(const ab_image_stats *st, const ab_auto_stf_config *cfg, ab_stf_params *out) {
    if (!st || !cfg || !out) return AB_ERR_INVALID;
    out->shadow = 0.0;
    out->midtone = 0.5;
    out->highlight = 1.0;
    if (st->valid_count == 0) return AB_OK;
    auto clampd = [](double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); };
    const double range = std::fmax(st->max - st->min, 1e-30);
    const double median_norm = (st->median - st->min) / range;
    const double sigma_norm = st->sigma / range;
    const double shadow_norm = clampd(median_norm + cfg->shadow_k * sigma_norm, 0.0, 0.98);
    const double highlight_norm = 1.0;
    const double clip_range = std::fmax(highlight_norm - shadow_norm, 1e-15);
    const double m_clipped = clampd((median_norm - shadow_norm) / clip_range, 0.0, 1.0);
    double midtone;
    if (m_clipped <= 0.0 || m_clipped >= 1.0) {
        midtone = 0.5;
    } else {
        const double m = m_clipped, t = cfg->target_bg;
        const double denom = 2.0 * t * m - t - m;
        midtone = std::fabs(denom) < 1e-15 ? 0.5 : clampd(m * (t - 1.0) / denom, 0.0001, 0.9999);
    }
    out->shadow = shadow_norm;
    out->midtone = midtone;
    out->highlight = highlight_norm;
    return AB_OK;
}

// linked-STF stat merge, cmd/helpers.rs:185-202
ab_status ab_linked_stats(const ab_image_stats *r, const ab_image_stats *g, const ab_image_stats *b,
                          ab_image_stats *c) {
    if (!r || !g || !b || !c) return AB_ERR_INVALID;
    c->min = std::fmin(std::fmin(r->min, g->min), b->min);
    c->max = std::fmax(std::fmax(r->max, g->max), b->max);
    c->mean = (r->mean + g->mean + b->mean) / 3.0;
    c->median = (r->median + g->median + b->median) / 3.0;
    c->sigma = std::sqrt((r->sigma * r->sigma + g->sigma * g->sigma + b->sigma * b->sigma) / 3.0);
    c->mad = (r->mad + g->mad + b->mad) / 3.0;
    c->valid_count = r->valid_count;
    return AB_OK;
}

// select_wb_reference, core/compose/white_balance.rs:3-20
ab_status ab_select_wb_reference(const ab_image_stats *sr, const ab_image_stats *sg,
                                 const ab_image_stats *sb, double out3[3]) {
    if (!sr || !sg || !sb || !out3) return AB_ERR_INVALID;
    auto stab = [](const ab_image_stats *s) { return s->median > 1e-10 ? s->mad / s->median : DBL_MAX; };
    const double a = stab(sr), b = stab(sg), c = stab(sb);
    if (a <= b && a <= c) {
        const double m = std::fmax(sr->median, 1e-10);
        out3[0] = 1.0;
        out3[1] = m / std::fmax(sg->median, 1e-10);
        out3[2] = m / std::fmax(sb->median, 1e-10);
    } else if (c <= b) {
        const double m = std::fmax(sb->median, 1e-10);
        out3[0] = m / std::fmax(sr->median, 1e-10);
        out3[1] = m / std::fmax(sg->median, 1e-10);
        out3[2] = 1.0;
    } else {
        const double m = std::fmax(sg->median, 1e-10);
        out3[0] = m / std::fmax(sr->median, 1e-10);
        out3[1] = 1.0;
        out3[2] = m / std::fmax(sb->median, 1e-10);
    }
    return AB_OK;
}

}  // extern "C"
