// stack_kernels.cuh -- device code of the sigma-clipped frame stack (sm_100a).
//
// Reference semantics: stack_images / sigma_clip_combine,
// src-tauri/src/core/stacking/combine.rs:14-92,160-182 (align=false core loop).
//
// Kernel structure (one thread = one output pixel, P = exact frame count, samples in registers)
//
//  phase A  every pixel, straight-line, register-resident:
//    * P coalesced streaming loads (one 128 B line per warp per frame plane), or P conflict-free
//      LDS from a TMA-staged tile (kTma variant, see stack_tile_kernel);
//    * Batcher odd-even merge sorting network (191 compare-exchanges at P=32, FMNMX pairs);
//      non-finite samples map to +inf and sort last (combine.rs:170-175 drops them);
//    * iteration 0 (combine.rs:37-48) from the sorted order: med = s[P/2] (upper median) and,
//      because |v-med| is V-shaped over sorted samples, the upper-median MAD is
//      min_t max(med - s[t], s[t+P/2] - med);
//    * the clip (combine.rs:65-82): dev = v - center is monotone in v, so the rejects are a
//      prefix and a suffix of the sorted samples -- peeled from both ends with early exit;
//    * ~80 % of pixels reject nothing at iteration 0, the reference loop then stops
//      (combine.rs:80-82) and the result is the f64 mean of all P samples.
//  phase B  only pixels that rejected something (or have non-finite samples): their sorted
//    samples are parked in shared memory, compacted over the block, and finished by the first
//    `count` threads with the dynamic-index version of the same algorithm (mean / sample
//    variance in f64 over the surviving run, peel, repeat).  This keeps the rare, divergent
//    iterations >= 1 from being executed in lock-step by every warp.
//
// Arithmetic mirrors the reference expression by expression (f32 compares, f64 sums, casts,
// 1e-10 floors); compiled with -fmad=false.  The f64 sums run in a fixed 4-way interleaved
// ascending order; the reference sums in select_nth-permuted order, identical whenever the f64
// sum of f32 samples is exact (dynamic range < 2^29).
#pragma once

#include <cstdint>
#include <utility>

#include <math_constants.h>

#include "ab_internal.h"

namespace abk {

constexpr double MAD_TO_SIGMA = 1.4826;  // types/constants.rs:7
constexpr unsigned META_FINAL_ONLY = 0x80000000u;  // parked pixel: only the final mean is left
constexpr int TILE = 256;                // pixels per block iteration = threads per block

__device__ __forceinline__ void ce(float &a, float &b) {
    const float lo = fminf(a, b);
    const float hi = fmaxf(a, b);
    a = lo;
    b = hi;
}

// Batcher odd-even merge sort network for any compile-time P, generated as a constexpr list of
// compare-exchange pairs and expanded with a fold expression, so every index is a literal and
// the array provably stays in registers.
struct CePair {
    short a, b;
};
template <int P>
struct OemNet {
    static constexpr int kMax = P * 12 + 4;
    CePair pr[kMax];
    int count;
    constexpr OemNet() : pr{}, count(0) {
        int p0 = 1;
        if (P % 16 == 0) {
            // 16-blocks: the 60-comparator, 10-layer 16-input network (Green; verified over all
            // 2^16 zero-one inputs by tools/verify_network.py) instead of Batcher's 63
            constexpr short g[60][2] = {
                {0, 13}, {1, 12}, {2, 15}, {3, 14}, {4, 8},   {5, 6},   {7, 11},  {9, 10},
                {0, 5},  {1, 7},  {2, 9},  {3, 4},  {6, 13},  {8, 14},  {10, 15}, {11, 12},
                {0, 1},  {2, 3},  {4, 5},  {6, 8},  {7, 9},   {10, 11}, {12, 13}, {14, 15},
                {0, 2},  {1, 3},  {4, 10}, {5, 11}, {6, 7},   {8, 9},   {12, 14}, {13, 15},
                {1, 2},  {3, 12}, {4, 6},  {5, 7},  {8, 10},  {9, 11},  {13, 14},
                {1, 4},  {2, 6},  {5, 8},  {7, 10}, {9, 13},  {11, 14},
                {2, 4},  {3, 6},  {9, 12}, {11, 13},
                {3, 5},  {6, 8},  {7, 9},  {10, 12},
                {3, 4},  {5, 6},  {7, 8},  {9, 10}, {11, 12},
                {6, 7},  {8, 9}};
            for (int blk = 0; blk < P / 16; ++blk)
                for (int i = 0; i < 60; ++i) {
                    pr[count].a = (short)(blk * 16 + g[i][0]);
                    pr[count].b = (short)(blk * 16 + g[i][1]);
                    ++count;
                }
            p0 = 16;
        }
        for (int p = p0; p < P; p <<= 1)
            for (int k = p; k >= 1; k >>= 1)
                for (int j = k % p; j + k < P; j += 2 * k)
                    for (int i = 0; i < k; ++i)
                        if (i + j + k < P && (i + j) / (2 * p) == (i + j + k) / (2 * p)) {
                            pr[count].a = (short)(i + j);
                            pr[count].b = (short)(i + j + k);
                            ++count;
                        }
    }
};
template <int P>
struct OemHolder {
    static constexpr OemNet<P> net{};
};
template <int P, int... I>
__device__ __forceinline__ void sort_network_impl(float (&s)[P], std::integer_sequence<int, I...>) {
    (ce(s[OemHolder<P>::net.pr[I].a], s[OemHolder<P>::net.pr[I].b]), ...);
}
template <int P>
__device__ __forceinline__ void sort_network(float (&s)[P]) {
    if constexpr (P > 1)
        sort_network_impl<P>(s, std::make_integer_sequence<int, OemHolder<P>::net.count>{});
}

// ---- f64 sums of the surviving run s[a..b) of the ascending samples.
//
// The reference accumulates `v as f64` (combine.rs:51,90).  Converting every sample costs one
// F2F.F64.F32 each, and that instruction issues at 16 lanes/clk/SM (measured), so the conversions --
// not the DADDs -- dominate the mean.  When the run is positive and spans at most a factor of two
// (s[b-1] <= 2 s[a], the normal case for sky-dominated pixels) every difference s[j] - s[a] is
// exact in f32 (Sterbenz) and a multiple of ulp(s[a]); their non-negative partial sums stay exact
// in f32 as long as the total is below 2^24 ulps = 2 * 2^floor(log2 s[a]).  Then
//     sum = len * s[a] + D        (exact in f64: < 2^31 ulps)
// is the exact real sum, which is also what the reference's f64 accumulation yields for such a run,
// so the result is bit-identical.  The certificate is checked per pixel; anything else takes the
// plain f64 accumulation (4 accumulators keyed by the index relative to a, combined pairwise).
__device__ __forceinline__ float next_pow2_above(float s0) {  // 2^(floor(log2 s0) + 1), s0 normal > 0
    return __uint_as_float((__float_as_uint(s0) & 0x7f800000u) + 0x00800000u);
}
// Register version (all P samples survive).  Differences are taken around the median so the two
// one-signed partial sums stay small (a no-reject pixel has |s[j] - med| <= ~4.4 MAD); returns
// false when the certificate fails -- the caller then parks the pixel for the f64 path of phase B.
template <int P>
__device__ __forceinline__ bool mean_all_reg(const float (&s)[P], float &mean_out) {
    const float s0 = s[0], c = s[P / 2];
    float lo[2] = {0.0f, 0.0f}, hi[2] = {0.0f, 0.0f};
#pragma unroll
    for (int j = 0; j < P / 2; ++j) lo[j & 1] += s[j] - c;  // <= 0
#pragma unroll
    for (int j = P / 2 + 1; j < P; ++j) hi[j & 1] += s[j] - c;  // >= 0
    const float d_lo = lo[0] + lo[1], d_hi = hi[0] + hi[1];
    const float T = next_pow2_above(s0);
    const bool ok = s0 >= 1.17549435e-38f && s[P - 1] <= 2.0f * s0 && d_hi < T && -d_lo < T;
    const double sum = (double)P * (double)c + (double)(d_hi + d_lo);
    mean_out = (float)(sum / (double)P);
    return ok;
}
__device__ __forceinline__ double sum_run_mem(const float *s, int stride, int a, int b) {
    const float s0 = s[a * stride];
    if (s0 >= 1.17549435e-38f && s[(b - 1) * stride] <= 2.0f * s0) {
        const int mid = (a + b) >> 1;
        const float c = s[mid * stride];
        float lo = 0.0f, hi = 0.0f;
#pragma unroll 2
        for (int j = a; j < mid; ++j) lo += s[j * stride] - c;
#pragma unroll 2
        for (int j = mid + 1; j < b; ++j) hi += s[j * stride] - c;
        const float T = next_pow2_above(s0);
        if (hi < T && -lo < T) return (double)(b - a) * (double)c + (double)(hi + lo);
    }
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int j = a;
#pragma unroll 1
    for (; j + 4 <= b; j += 4) {
        acc0 += (double)s[j * stride];
        acc1 += (double)s[(j + 1) * stride];
        acc2 += (double)s[(j + 2) * stride];
        acc3 += (double)s[(j + 3) * stride];
    }
    if (j < b) acc0 += (double)s[j * stride];
    if (j + 1 < b) acc1 += (double)s[(j + 1) * stride];
    if (j + 2 < b) acc2 += (double)s[(j + 2) * stride];
    return (acc0 + acc1) + (acc2 + acc3);
}
__device__ __forceinline__ double sumsq_run_mem(const float *s, int stride, int a, int b,
                                                double mean) {
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int j = a;
#pragma unroll 1
    for (; j + 4 <= b; j += 4) {
        const double e0 = (double)s[j * stride] - mean, e1 = (double)s[(j + 1) * stride] - mean;
        const double e2 = (double)s[(j + 2) * stride] - mean, e3 = (double)s[(j + 3) * stride] - mean;
        acc0 += e0 * e0;
        acc1 += e1 * e1;
        acc2 += e2 * e2;
        acc3 += e3 * e3;
    }
    if (j < b) {
        const double e = (double)s[j * stride] - mean;
        acc0 += e * e;
    }
    if (j + 1 < b) {
        const double e = (double)s[(j + 1) * stride] - mean;
        acc1 += e * e;
    }
    if (j + 2 < b) {
        const double e = (double)s[(j + 2) * stride] - mean;
        acc2 += e * e;
    }
    return (acc0 + acc1) + (acc2 + acc3);
}

// Dynamic-index sigma_clip_combine on ascending samples s[0..n) (stride `stride`), resumed at
// iteration `it0` with the surviving run [a, b), `rejected` rejections so far and the last
// center.  it0 == 0 runs the whole algorithm (combine.rs:14-92).
__device__ __forceinline__ float clip_combine_sorted_mem(const float *s, int stride, int n, int a,
                                                         int b, unsigned it0, float last_center,
                                                         unsigned &rejected, float sigma_low,
                                                         float sigma_high, unsigned max_iter) {
    if (n == 0) return 0.0f;         // combine.rs:21-23
    if (n == 1) return s[0];         // combine.rs:24-26
    bool mean_known = false;
    double mean_cur = 0.0;
    for (unsigned it = it0; it < max_iter; ++it) {
        const int len = b - a;
        if (len < 2) break;
        float center, sigma;
        if (it == 0) {
            const int m = n >> 1;
            const float med = s[m * stride];
            float mad = CUDART_INF_F;
            for (int t = 0; t <= m; ++t) {
                const float left = med - s[t * stride];
                const float right = (t + m < n) ? s[(t + m) * stride] - med : CUDART_INF_F;
                mad = fminf(mad, fmaxf(left, right));
            }
            center = med;
            sigma = (float)fmax((double)mad * MAD_TO_SIGMA, 1e-10);
        } else {
            const double nn = (double)len;
            const double mean = sum_run_mem(s, stride, a, b) / nn;
            const double ss = sumsq_run_mem(s, stride, a, b, mean);
            center = (float)mean;
            sigma = (float)fmax(sqrt(ss / fmax(nn - 1.0, 1.0)), 1e-10);
            mean_known = true;
            mean_cur = mean;
        }
        last_center = center;
        const float lo = -sigma_low * sigma;
        const float hi = sigma_high * sigma;
        // dev = v - center is monotone over the ascending samples and never NaN for finite v:
        // the rejects are a prefix (dev < lo) and a suffix (dev > hi).  The negated compares make
        // NaN bounds reject everything, like the reference's `dev >= lo && dev <= hi`.
        int na = a, nb = b;
        while (na < nb && !(s[na * stride] - center >= lo)) ++na;
        while (nb > na && !(s[(nb - 1) * stride] - center <= hi)) --nb;
        const int removed = len - (nb - na);
        rejected += (unsigned)removed;
        a = na;
        b = nb;
        if (removed == 0) break;
        mean_known = false;
    }
    const int len = b - a;
    if (len == 0) return isfinite(last_center) ? last_center : 0.0f;  // combine.rs:85-88
    if (mean_known) return (float)mean_cur;  // same run, same summation order as the last pass
    return (float)(sum_run_mem(s, stride, a, b) / (double)len);
}

// meta word of a parked pixel: a | b << 8 | it0 << 16, bit 31 = only the final mean is left
__device__ __forceinline__ void decode_meta(unsigned meta, int P, unsigned &it0, int &n, int &a,
                                            int &b) {
    it0 = (meta >> 16) & 0x7fffu;
    if (meta & META_FINAL_ONLY) {
        n = P;
        a = 0;
        b = P;
        it0 = 0xffffffffu;  // >= max_iter: the clip loop is skipped
    } else if (it0 == 0) {
        n = (meta >> 8) & 0xff;
        a = 0;
        b = n;
    } else {
        n = P;
        a = meta & 0xff;
        b = (meta >> 8) & 0xff;
    }
}

// Phase A on a register array of exactly P finite-or-+inf samples (n of them finite).
// Returns true when `result`/`rejected` are final; false when the pixel must be finished by
// phase B, in which case s[] is sorted and meta = a | b << 8 | it0 << 16.
template <int P>
__device__ __forceinline__ bool phase_a(float (&s)[P], int n, float sigma_low, float sigma_high,
                                        unsigned max_iter, float &result, unsigned &rejected,
                                        unsigned &meta) {
    rejected = 0;
    sort_network<P>(s);
    if (n != P) {  // non-finite samples present: rare, finished with dynamic indices
        if (n == 0) {
            result = 0.0f;
            return true;
        }
        if (n == 1) {
            result = s[0];
            return true;
        }
        meta = 0u | ((unsigned)n << 8) | (0u << 16);
        return false;
    }
    if constexpr (P == 1) {
        result = s[0];
        return true;
    } else {
        if (max_iter == 0) {
            if (mean_all_reg<P>(s, result)) return true;
            meta = 0u | ((unsigned)P << 8) | META_FINAL_ONLY;
            result = 0.0f;
            return false;
        }
        // iteration 0: upper median, upper-median MAD (combine.rs:37-48)
        const float med = s[P / 2];
        float mad = CUDART_INF_F;
#pragma unroll
        for (int t = 0; t < (P + 1) / 2; ++t)
            mad = fminf(mad, fmaxf(med - s[t], s[t + P / 2] - med));
        // (for even P the last window pairs s[P/2] with the absent s[P] = +inf: skipped)
        const float sigma = (float)fmax((double)mad * MAD_TO_SIGMA, 1e-10);
        const float lo = -sigma_low * sigma;
        const float hi = sigma_high * sigma;
        // peel the rejects off both ends; checked in groups of 4 so the common case (nothing or
        // one sample rejected) leaves after the first group
        int nh = 0, nl = 0;
        {
            bool go = true;
#pragma unroll
            for (int j = P - 1; j >= 0; --j) {
                const bool fail = go && !(s[j] - med <= hi);
                nh += fail ? 1 : 0;
                go = fail;
                if (((P - 1 - j) & 3) == 3 && !go) break;
            }
            go = true;
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const bool fail = go && !(s[j] - med >= lo);
                nl += fail ? 1 : 0;
                go = fail;
                if ((j & 3) == 3 && !go) break;
            }
        }
        if (nl + nh == 0) {  // removed == 0 -> loop ends (combine.rs:80-82); mean of everything
            if (mean_all_reg<P>(s, result)) return true;
            meta = 0u | ((unsigned)P << 8) | META_FINAL_ONLY;  // rare: f64 accumulation in phase B
            result = med;
            return false;
        }
        if (nl + nh >= P) {  // everything rejected: last center (combine.rs:85-88)
            rejected = P;
            result = isfinite(med) ? med : 0.0f;
            return true;
        }
        rejected = (unsigned)(nl + nh);
        meta = (unsigned)nl | ((unsigned)(P - nh) << 8) | (1u << 16);
        result = med;  // last_center carried into phase B
        return false;
    }
}

// smem layout per block: col[P][TILE] sorted samples parked by phase A (column = thread),
// list[TILE] compacted thread ids, meta[TILE], lastc[TILE], count.
template <int P>
struct StackSmem {
    float col[P][TILE];
    float lastc[TILE];
    unsigned meta[TILE];
    unsigned rej0[TILE];
    unsigned short list[TILE];
    unsigned count[2];
    unsigned long long rej_total;
};

template <int P>
__global__ void __launch_bounds__(TILE)
stack_tile_kernel(const __grid_constant__ FrameTable ft, size_t npix, float sigma_low,
                  float sigma_high, unsigned max_iter, float *__restrict__ out,
                  unsigned long long *__restrict__ rejected_total) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StackSmem<P> &sm = *reinterpret_cast<StackSmem<P> *>(smem_raw);
    const int tid = threadIdx.x;
    if (tid == 0) {
        sm.count[0] = 0;
        sm.count[1] = 0;
        sm.rej_total = 0;
    }
    __syncthreads();
    unsigned local_rej = 0;
    unsigned par = 0;
    const size_t n_tiles = (npix + TILE - 1) / TILE;
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, par ^= 1u) {
        const size_t p = tile * TILE + tid;
        const bool active = p < npix;
        if (active) {
            float s[P];
            float probe = 0.0f;
#pragma unroll
            for (int f = 0; f < P; ++f) {
                s[f] = __ldcs(ft.p[f] + p);
                probe = fmaf(s[f], 0.0f, probe);  // NaN iff any sample is non-finite
            }
            int n = P;
            if (probe != probe) {
                n = 0;
#pragma unroll
                for (int f = 0; f < P; ++f) {
                    const bool fin = fabsf(s[f]) < CUDART_INF_F;
                    n += fin ? 1 : 0;
                    s[f] = fin ? s[f] : CUDART_INF_F;
                }
            }
            float result;
            unsigned rej, meta;
            if (phase_a<P>(s, n, sigma_low, sigma_high, max_iter, result, rej, meta)) {
                __stcs(out + p, result);
                local_rej += rej;
            } else {
#pragma unroll
                for (int f = 0; f < P; ++f) sm.col[f][tid] = s[f];
                sm.meta[tid] = meta;
                sm.lastc[tid] = result;
                sm.rej0[tid] = rej;
                const unsigned slot = atomicAdd(&sm.count[par], 1u);
                sm.list[slot] = (unsigned short)tid;
            }
        }
        __syncthreads();
        const unsigned cnt = sm.count[par];
        if (tid == 0) sm.count[par ^ 1u] = 0;  // last read before the previous tile's 2nd barrier
        if ((unsigned)tid < cnt) {
            const int src = sm.list[tid];
            const unsigned meta = sm.meta[src];
            unsigned it0;
            int a, b, n;
            decode_meta(meta, P, it0, n, a, b);
            unsigned rej = sm.rej0[src];
            const float r = clip_combine_sorted_mem(&sm.col[0][src], TILE, n, a, b, it0,
                                                    sm.lastc[src], rej, sigma_low, sigma_high,
                                                    max_iter);
            __stcs(out + tile * TILE + src, r);
            local_rej += rej;
        }
        __syncthreads();  // phase B done: columns and the list may be overwritten
    }
    const unsigned w = __reduce_add_sync(0xffffffffu, local_rej);
    if ((tid & 31) == 0 && w) atomicAdd(&sm.rej_total, (unsigned long long)w);
    __syncthreads();
    if (tid == 0 && sm.rej_total) atomicAdd(rejected_total, sm.rej_total);
}

// ---------------------------------------------------------------------------------------------
// Warp-autonomous variant: no block-wide synchronisation at all.
//
// Every warp walks its own sequence of 32-pixel groups.  Lanes whose pixel needs phase B park the
// sorted samples in a warp-private ring in shared memory (column = ring slot, conflict-free); as
// soon as the ring holds >= 32 parked pixels the warp finishes 32 of them at once, one per lane,
// so phase B always runs with full lanes and only every ~6th group.  While one warp is in its
// (FP64 / XU heavy) phase B the other warps of the SM keep the ALU pipe busy with their sorting
// networks: warps drift apart instead of marching in lock-step, which is what keeps the ALU pipe
// (the real bound of this kernel) fed.
constexpr int WQ_CAP = 64;         // ring slots per warp (power of two)
constexpr int WARPS_PER_BLOCK = 8;

template <int P, int CAP = WQ_CAP>
struct WarpQueue {
    static constexpr int kCap = CAP;
    float v[P][CAP];
    float lastc[CAP];
    unsigned meta[CAP];
    unsigned rej0[CAP];
    unsigned pix[CAP];  // pixel index relative to the launch (npix < 2^32 per launch)
};

// finish `cnt` (<= 32) parked pixels starting at ring position head; lane i takes entry i
template <int P, int CAP>
__device__ __forceinline__ void drain_batch(WarpQueue<P, CAP> &q, unsigned head, unsigned cnt, int lane,
                                            float sigma_low, float sigma_high, unsigned max_iter,
                                            float *__restrict__ out, unsigned &local_rej) {
    if ((unsigned)lane < cnt) {
        const unsigned slot = (head + lane) & (CAP - 1);
        const unsigned meta = q.meta[slot];
        unsigned it0;
        int a, b, n;
        decode_meta(meta, P, it0, n, a, b);
        unsigned rej = q.rej0[slot];
        const float r = clip_combine_sorted_mem(&q.v[0][slot], CAP, n, a, b, it0, q.lastc[slot],
                                                rej, sigma_low, sigma_high, max_iter);
        __stcs(out + q.pix[slot], r);
        local_rej += rej;
    }
    __syncwarp();
}

template <int P, int kMinBlocks, bool kPrefetch>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, kMinBlocks)
stack_warp_kernel(const __grid_constant__ FrameTable ft, size_t npix, float sigma_low,
                  float sigma_high, unsigned max_iter, float *__restrict__ out,
                  unsigned long long *__restrict__ rejected_total) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpQueue<P> *queues = reinterpret_cast<WarpQueue<P> *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpQueue<P> &q = queues[warp];
    unsigned head = 0, count = 0;  // warp-uniform ring state
    unsigned local_rej = 0;

    const size_t n_groups = (npix + 31) / 32;
    const size_t warp_global = (size_t)blockIdx.x * WARPS_PER_BLOCK + warp;
    const size_t warp_stride = (size_t)gridDim.x * WARPS_PER_BLOCK;
    float nxt[kPrefetch ? P : 1];
    if constexpr (kPrefetch) {
        const size_t p0 = warp_global * 32 + lane;
        if (p0 < npix) {
#pragma unroll
            for (int f = 0; f < P; ++f) nxt[f] = __ldcs(ft.p[f] + p0);
        }
    }
    for (size_t g = warp_global; g < n_groups; g += warp_stride) {
        const size_t p = g * 32 + lane;
        const bool active = p < npix;
        bool park = false;
        float s[P];
        float result = 0.0f;
        unsigned rej = 0, meta = 0;
        if constexpr (kPrefetch) {
#pragma unroll
            for (int f = 0; f < P; ++f) s[f] = nxt[f];
            const size_t pn = p + warp_stride * 32;  // software pipeline: next group's loads in flight
            if (pn < npix) {
#pragma unroll
                for (int f = 0; f < P; ++f) nxt[f] = __ldcs(ft.p[f] + pn);
            }
        }
        if (active) {
            float probe = 0.0f;
#pragma unroll
            for (int f = 0; f < P; ++f) {
                if constexpr (!kPrefetch) s[f] = __ldcs(ft.p[f] + p);
                probe = fmaf(s[f], 0.0f, probe);  // NaN iff any sample is non-finite
            }
            int n = P;
            if (probe != probe) {
                n = 0;
#pragma unroll
                for (int f = 0; f < P; ++f) {
                    const bool fin = fabsf(s[f]) < CUDART_INF_F;
                    n += fin ? 1 : 0;
                    s[f] = fin ? s[f] : CUDART_INF_F;
                }
            }
            if (phase_a<P>(s, n, sigma_low, sigma_high, max_iter, result, rej, meta)) {
                __stcs(out + p, result);
                local_rej += rej;
            } else {
                park = true;
            }
        }
        const unsigned pmask = __ballot_sync(0xffffffffu, park);
        if (pmask) {
            // count < 32 here and at most 32 lanes park, so the 64-slot ring always has room
            if (park) {
                const unsigned slot =
                    (head + count + __popc(pmask & ((1u << lane) - 1u))) & (WQ_CAP - 1);
#pragma unroll
                for (int f = 0; f < P; ++f) q.v[f][slot] = s[f];
                q.meta[slot] = meta;
                q.lastc[slot] = result;
                q.rej0[slot] = rej;
                q.pix[slot] = (unsigned)p;
            }
            count += __popc(pmask);
            __syncwarp();
            if (count >= 32u) {
                drain_batch<P, WQ_CAP>(q, head, 32u, lane, sigma_low, sigma_high, max_iter, out, local_rej);
                head = (head + 32u) & (WQ_CAP - 1);
                count -= 32u;
            }
        }
    }
    if (count) drain_batch<P, WQ_CAP>(q, head, count, lane, sigma_low, sigma_high, max_iter, out, local_rej);
    const unsigned w = __reduce_add_sync(0xffffffffu, local_rej);
    if (lane == 0 && w) atomicAdd(rejected_total, (unsigned long long)w);
}

// ---------------------------------------------------------------------------------------------
// TMA-staged, warp-specialised variant (the production path for 16-byte aligned frames).
//
//   * a tile = 256 consecutive pixels of all P frames.  When the frames sit at a constant stride
//     (one [P, rows, cols] cube: the layout the host-buffer entry point and bench.py use) the
//     producer warp fetches a tile with ONE 2-D tensor-map TMA (cp.async.bulk.tensor.2d, box
//     256 x P, UTMALDG); otherwise with P 1-D bulk copies (cp.async.bulk, UBLKCP).  Completion is
//     signalled on an mbarrier (complete_tx::bytes); stages form a ring;
//   * 8 "A" warps run phase A on a landed stage: conflict-free LDS with immediate offsets (no
//     per-load address arithmetic on the ALU pipe, which the sorting network saturates), then
//     release the stage at once (the samples are in registers), sort, do iteration 0 and either
//     store the pixel or park the sorted samples in a small double-buffered queue;
//   * 2 "B" warps finish the parked pixels of tile k with the dynamic-index code while the A warps
//     are already sorting tile k+1.  No block-wide barrier anywhere, only mbarriers;
//   * a parked pixel that finds the queue full (NaN strips: every pixel of the tile parks) is
//     finished by its own thread from a local-memory copy.
constexpr int TMA_STAGES = 2;
constexpr int A_THREADS = TILE;  // 8 warps
constexpr int B_THREADS = 64;    // 2 warps
constexpr int PROD_THREADS = 32; // 1 warp, one elected lane
constexpr int QCAP = 96;         // parked pixels per tile before the local-memory fallback

template <int P>
struct StackTmaSmem {
    alignas(128) float stage[TMA_STAGES][P][TILE];
    float q[2][P][QCAP];          // parked sorted samples, column = queue slot
    float q_lastc[2][QCAP];
    unsigned q_meta[2][QCAP];     // a | b << 8 | it0 << 16
    unsigned q_rej0[2][QCAP];
    unsigned short q_src[2][QCAP];
    unsigned q_count[2];
    alignas(8) unsigned long long full_bar[TMA_STAGES];   // TMA bytes landed
    alignas(8) unsigned long long empty_bar[TMA_STAGES];  // all A warps copied the stage to regs
    alignas(8) unsigned long long adone_bar[2];           // all A warps finished tile (parity)
    alignas(8) unsigned long long qfree_bar[2];           // B warps drained queue (parity)
    unsigned long long rej_total;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Latency-critical wait (A warps: data normally already landed).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// Wait for the helper roles (B warps, producer): a short sleep between polls keeps the spin from
// taking issue slots from the A warps.  (A large try_wait suspend-time hint compiles to
// NANOSLEEP.SYNCS and was measured to oversleep the ~2 us tile period: the TMA ring starved.)
template <int NS>
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(NS);
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst_smem, const void *src, uint32_t bytes,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(dst_smem),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const void *tmap, int x, int y,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
        "l"(tmap), "r"(x), "r"(y), "r"(bar)
        : "memory");
}

// Out-of-line so the B warps and the (rare) overflow path share one copy of the code.
static __device__ __noinline__ float finish_parked(const float *s, int stride, unsigned meta, int P,
                                            float last_center, unsigned &rej, float sigma_low,
                                            float sigma_high, unsigned max_iter) {
    unsigned it0;
    int a, b, n;
    decode_meta(meta, P, it0, n, a, b);
    return clip_combine_sorted_mem(s, stride, n, a, b, it0, last_center, rej, sigma_low,
                                   sigma_high, max_iter);
}

struct TmapParam {
    alignas(64) unsigned char bytes[128];  // CUtensorMap
};

template <int P, bool kTensorMap>
__global__ void __launch_bounds__(A_THREADS + B_THREADS + PROD_THREADS)
stack_tma_kernel(const __grid_constant__ FrameTable ft, const __grid_constant__ TmapParam tmap,
                 size_t n_tiles, float sigma_low, float sigma_high, unsigned max_iter,
                 float *__restrict__ out, unsigned long long *__restrict__ rejected_total,
                 unsigned dbg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StackTmaSmem<P> &sm = *reinterpret_cast<StackTmaSmem<P> *>(smem_raw);
    const int tid = threadIdx.x;
    constexpr uint32_t SEG_BYTES = TILE * sizeof(float);
    constexpr uint32_t TILE_BYTES = P * SEG_BYTES;

    if (tid == 0) {
        for (int st = 0; st < TMA_STAGES; ++st) {
            mbar_init(smem_u32(&sm.full_bar[st]), 1);
            mbar_init(smem_u32(&sm.empty_bar[st]), A_THREADS / 32);
        }
        for (int q = 0; q < 2; ++q) {
            mbar_init(smem_u32(&sm.adone_bar[q]), A_THREADS / 32);
            mbar_init(smem_u32(&sm.qfree_bar[q]), B_THREADS / 32);
            sm.q_count[q] = 0;
        }
        sm.rej_total = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // tiles of this block: blockIdx.x, blockIdx.x + gridDim.x, ...
    const size_t my_tiles =
        (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    unsigned local_rej = 0;

    if (tid >= A_THREADS + B_THREADS) {
        // ---------------------------------------------------------------- producer (one lane)
        if (tid == A_THREADS + B_THREADS) {
            for (size_t k = 0; k < my_tiles; ++k) {
                const int st = (int)(k % TMA_STAGES);
                if (k >= TMA_STAGES)
                    mbar_wait_sleep<20>(smem_u32(&sm.empty_bar[st]),
                                    (uint32_t)(((k / TMA_STAGES) - 1) & 1));
                const size_t tile = blockIdx.x + k * gridDim.x;
                const uint32_t bar = smem_u32(&sm.full_bar[st]);
                mbar_arrive_expect_tx(bar, TILE_BYTES);
                if constexpr (kTensorMap) {
                    tma_load_2d(smem_u32(&sm.stage[st][0][0]), &tmap, (int)(tile * TILE), 0, bar);
                } else {
#pragma unroll 4
                    for (int f = 0; f < P; ++f)
                        tma_load_1d(smem_u32(&sm.stage[st][f][0]), ft.p[f] + tile * TILE, SEG_BYTES,
                                    bar);
                }
            }
        }
    } else if (tid >= A_THREADS) {
        // ---------------------------------------------------------------- B warps
        const int bt = tid - A_THREADS;
        for (size_t k = 0; k < my_tiles; ++k) {
            const int qb = (int)(k & 1);
            mbar_wait_sleep<100>(smem_u32(&sm.adone_bar[qb]), (uint32_t)((k >> 1) & 1));
            unsigned cnt = sm.q_count[qb];
            if (cnt > QCAP) cnt = QCAP;
            const size_t tile = blockIdx.x + k * gridDim.x;
            if (dbg & 2u) cnt = 0;  // timing experiment: B warps skip their work
            for (unsigned e = bt; e < cnt; e += B_THREADS) {
                unsigned rej = sm.q_rej0[qb][e];
                const float r = finish_parked(&sm.q[qb][0][e], QCAP, sm.q_meta[qb][e], P,
                                              sm.q_lastc[qb][e], rej, sigma_low, sigma_high,
                                              max_iter);
                __stcs(out + tile * TILE + sm.q_src[qb][e], r);
                local_rej += rej;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(B_THREADS) : "memory");
            if (bt == 0) sm.q_count[qb] = 0;
            __syncwarp();
            if ((bt & 31) == 0) mbar_arrive(smem_u32(&sm.qfree_bar[qb]));
        }
    } else {
        // ---------------------------------------------------------------- A warps
        for (size_t k = 0; k < my_tiles; ++k) {
            const int st = (int)(k % TMA_STAGES);
            const int qb = (int)(k & 1);
            const size_t tile = blockIdx.x + k * gridDim.x;
            mbar_wait(smem_u32(&sm.full_bar[st]), (uint32_t)((k / TMA_STAGES) & 1));
            float s[P];
            float probe = 0.0f;
#pragma unroll
            for (int f = 0; f < P; ++f) {
                s[f] = sm.stage[st][f][tid];
                probe = fmaf(s[f], 0.0f, probe);  // NaN iff any sample is non-finite
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(smem_u32(&sm.empty_bar[st]));  // stage may be refilled
            int n = P;
            if (probe != probe) {
                n = 0;
#pragma unroll
                for (int f = 0; f < P; ++f) {
                    const bool fin = fabsf(s[f]) < CUDART_INF_F;
                    n += fin ? 1 : 0;
                    s[f] = fin ? s[f] : CUDART_INF_F;
                }
            }
            float result;
            unsigned rej, meta;
            if (dbg & 4u) {  // timing experiment: load pipeline only, no sort / clip
                __stcs(out + tile * TILE + tid, probe + s[0] + s[P - 1]);
            } else if (phase_a<P>(s, n, sigma_low, sigma_high, max_iter, result, rej, meta)) {
                __stcs(out + tile * TILE + tid, result);
                local_rej += rej;
            } else {
                // the queue half is free once the B warps drained tile k-2
                if (k >= 2) mbar_wait(smem_u32(&sm.qfree_bar[qb]), (uint32_t)(((k >> 1) - 1) & 1));
                const unsigned slot = atomicAdd(&sm.q_count[qb], 1u);
                if (slot < QCAP) {
#pragma unroll
                    for (int f = 0; f < P; ++f) sm.q[qb][f][slot] = s[f];
                    sm.q_meta[qb][slot] = meta;
                    sm.q_lastc[qb][slot] = result;
                    sm.q_rej0[qb][slot] = rej;
                    sm.q_src[qb][slot] = (unsigned short)tid;
                } else {  // queue full: finish from a thread-private copy
                    float loc[P];
#pragma unroll
                    for (int f = 0; f < P; ++f) loc[f] = s[f];
                    const float r = finish_parked(loc, 1, meta, P, result, rej, sigma_low,
                                                  sigma_high, max_iter);
                    __stcs(out + tile * TILE + tid, r);
                    local_rej += rej;
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(smem_u32(&sm.adone_bar[qb]));
        }
    }
    const unsigned w = __reduce_add_sync(0xffffffffu, local_rej);
    if ((tid & 31) == 0 && w) atomicAdd(&sm.rej_total, (unsigned long long)w);
    __syncthreads();
    if (tid == 0 && sm.rej_total) atomicAdd(rejected_total, sm.rej_total);
}

// ---------------------------------------------------------------------------------------------
// Warp-autonomous variant with TMA staging (the production path for 16-byte aligned frames).
//
// Same control structure as stack_warp_kernel, but the 32 x P samples of a warp's next pixel group
// are fetched by the TMA engine into a warp-private 4 KiB shared-memory stage while the warp sorts
// the current group out of registers:
//   * frames at a constant stride (one [P, rows, cols] cube): ONE cp.async.bulk.tensor.2d per
//     group (box = 32 pixels x P frames, UTMALDG); otherwise P 128-byte cp.async.bulk copies;
//   * completion on a warp-private mbarrier (complete_tx::bytes); the stage is re-armed by lane 0
//     right after the warp has copied it to registers, so one stage per warp is enough: the fetch
//     of group i+1 overlaps the whole sort of group i;
//   * the loads become conflict-free LDS with immediate offsets: the ~1.5 ALU-pipe instructions
//     of 64-bit address arithmetic per global load disappear from the pipe that bounds the kernel.
constexpr int WT_WARPS = 8;   // 8 warps x ~8.6 KiB (P=32) -> 3 blocks / SM = 24 warps
constexpr int WT_CAP = 32;    // ring slots per warp; a group that does not fit drains first

template <int P>
struct WarpTmaSmem {
    alignas(128) float stage[P][32];
    WarpQueue<P, WT_CAP> q;
    alignas(8) unsigned long long full;
    unsigned pad[30];
};

template <int P, bool kTensorMap>
__global__ void __launch_bounds__(WT_WARPS * 32, 3)
stack_wtma_kernel(const __grid_constant__ FrameTable ft, const __grid_constant__ TmapParam tmap,
                  size_t n_groups, float sigma_low, float sigma_high, unsigned max_iter,
                  float *__restrict__ out, unsigned long long *__restrict__ rejected_total) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpTmaSmem<P> &sm = reinterpret_cast<WarpTmaSmem<P> *>(smem_raw)[warp];
    WarpQueue<P, WT_CAP> &q = sm.q;
    const uint32_t bar = smem_u32(&sm.full);
    const uint32_t stage_addr = smem_u32(&sm.stage[0][0]);
    constexpr uint32_t GROUP_BYTES = P * 32 * sizeof(float);

    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const size_t warp_global = (size_t)blockIdx.x * WT_WARPS + warp;
    const size_t warp_stride = (size_t)gridDim.x * WT_WARPS;

    auto issue = [&](size_t g) {  // lane 0 only
        mbar_arrive_expect_tx(bar, GROUP_BYTES);
        if constexpr (kTensorMap) {
            tma_load_2d(stage_addr, &tmap, (int)(g * 32), 0, bar);
        } else {
#pragma unroll 4
            for (int f = 0; f < P; ++f)
                tma_load_1d(stage_addr + f * 128, ft.p[f] + g * 32, 128, bar);
        }
    };
    if (lane == 0 && warp_global < n_groups) issue(warp_global);

    unsigned head = 0, count = 0;  // warp-uniform ring state
    unsigned local_rej = 0;
    uint32_t parity = 0;
    for (size_t g = warp_global; g < n_groups; g += warp_stride, parity ^= 1u) {
        const size_t p = g * 32 + lane;
        mbar_wait(bar, parity);
        float s[P];
        float probe = 0.0f;
#pragma unroll
        for (int f = 0; f < P; ++f) {
            s[f] = sm.stage[f][lane];
            probe = fmaf(s[f], 0.0f, probe);  // NaN iff any sample is non-finite
        }
        // every lane has consumed its column (probe depends on all 32 loads): refill the stage
        asm volatile("" ::"f"(probe) : "memory");
        __syncwarp();
        if (lane == 0 && g + warp_stride < n_groups) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(g + warp_stride);
        }
        int n = P;
        if (probe != probe) {
            n = 0;
#pragma unroll
            for (int f = 0; f < P; ++f) {
                const bool fin = fabsf(s[f]) < CUDART_INF_F;
                n += fin ? 1 : 0;
                s[f] = fin ? s[f] : CUDART_INF_F;
            }
        }
        float result;
        unsigned rej, meta;
        bool park = false;
        if (phase_a<P>(s, n, sigma_low, sigma_high, max_iter, result, rej, meta)) {
            __stcs(out + p, result);
            local_rej += rej;
        } else {
            park = true;
        }
        const unsigned pmask = __ballot_sync(0xffffffffu, park);
        if (pmask) {
            const unsigned c = __popc(pmask);
            if (count + c > (unsigned)WT_CAP) {  // make room: finish what is parked (< 32 lanes)
                drain_batch<P, WT_CAP>(q, head, count, lane, sigma_low, sigma_high, max_iter, out,
                                       local_rej);
                head = (head + count) & (WT_CAP - 1);
                count = 0;
            }
            if (park) {
                const unsigned slot =
                    (head + count + __popc(pmask & ((1u << lane) - 1u))) & (WT_CAP - 1);
#pragma unroll
                for (int f = 0; f < P; ++f) q.v[f][slot] = s[f];
                q.meta[slot] = meta;
                q.lastc[slot] = result;
                q.rej0[slot] = rej;
                q.pix[slot] = (unsigned)p;
            }
            count += c;
            __syncwarp();
            if (count >= 32u) {
                drain_batch<P, WT_CAP>(q, head, 32u, lane, sigma_low, sigma_high, max_iter, out,
                                       local_rej);
                head = (head + 32u) & (WT_CAP - 1);
                count -= 32u;
            }
        }
    }
    if (count)
        drain_batch<P, WT_CAP>(q, head, count, lane, sigma_low, sigma_high, max_iter, out, local_rej);
    const unsigned w = __reduce_add_sync(0xffffffffu, local_rej);
    if (lane == 0 && w) atomicAdd(rejected_total, (unsigned long long)w);
}

template <int P>
ab_status launch_ldg(ab_ctx *ctx, const FrameTable &ft, size_t npix, float sl, float sh,
                     unsigned max_iter, float *d_out, unsigned long long *d_rej,
                     cudaStream_t stream) {
    if (ctx->stack_variant == 1) {  // block-barrier phase B (kept for A/B measurements)
        const size_t smem = sizeof(StackSmem<P>);
        AB_CUDA(ctx, cudaFuncSetAttribute(stack_tile_kernel<P>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        AB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stack_tile_kernel<P>,
                                                                   TILE, smem));
        if (per_sm < 1) per_sm = 1;
        const size_t n_tiles = (npix + TILE - 1) / TILE;
        size_t grid = (size_t)ctx->sm_count * per_sm;
        if (grid > n_tiles) grid = n_tiles;
        stack_tile_kernel<P><<<(unsigned)grid, TILE, smem, stream>>>(ft, npix, sl, sh, max_iter,
                                                                    d_out, d_rej);
        AB_LAUNCH_CHECK(ctx);
        return AB_OK;
    }
    if (npix >= ((size_t)1 << 32))
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "npix %zu >= 2^32 per launch", npix);
    const size_t smem = sizeof(WarpQueue<P>) * WARPS_PER_BLOCK;
    constexpr int threads = WARPS_PER_BLOCK * 32;
    void (*kern)(FrameTable, size_t, float, float, unsigned, float *, unsigned long long *) =
        stack_warp_kernel<P, 3, false>;
    if (ctx->stack_variant == 4) kern = stack_warp_kernel<P, 4, false>;
    if (ctx->stack_variant == 5) kern = stack_warp_kernel<P, 2, true>;
    AB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    AB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) per_sm = 1;
    const size_t n_blocks_needed = (npix + threads - 1) / threads;
    size_t grid = (size_t)ctx->sm_count * per_sm;  // persistent: one wave, groups strided
    if (grid > n_blocks_needed) grid = n_blocks_needed;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, threads, smem, stream>>>(ft, npix, sl, sh, max_iter, d_out, d_rej);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

// Host side of the tensor map (stack.cu): encodes a [P, npix] f32 view with a 256 x P box.
bool ab_make_stack_tmap(ab_ctx *ctx, const float *base, size_t npix, size_t plane_stride_elems,
                        int P, int box_px, TmapParam *out);

template <int P>
ab_status launch_tile(ab_ctx *ctx, const FrameTable &ft, size_t npix, float sl, float sh,
                      unsigned max_iter, float *d_out, unsigned long long *d_rej,
                      cudaStream_t stream) {
    if (ctx->stack_variant == 0) {
        bool al = true;
        for (int f = 0; f < P; ++f) al = al && ((reinterpret_cast<uintptr_t>(ft.p[f]) & 15) == 0);
        const size_t groups = al && npix < ((size_t)1 << 31) ? npix / 32 : 0;
        if (groups) {
            bool cube = P > 1 && !ctx->force_tma1d;
            const ptrdiff_t stride = P > 1 ? ft.p[1] - ft.p[0] : 0;
            for (int f = 1; cube && f < P; ++f) cube = (ft.p[f] - ft.p[f - 1]) == stride;
            cube = cube && stride > 0 && (stride % 4) == 0 && (size_t)stride >= npix;
            TmapParam tm{};
            if (cube) cube = ab_make_stack_tmap(ctx, ft.p[0], npix, (size_t)stride, P, 32, &tm);
            // scattered frames: P small bulk copies per group measured slower (1.18 ms) than plain
            // coalesced loads (0.92 ms) -> only the tensor-map form is used unless forced
            if (!cube && !ctx->force_tma1d)
                return launch_ldg<P>(ctx, ft, npix, sl, sh, max_iter, d_out, d_rej, stream);
            const size_t smem = sizeof(WarpTmaSmem<P>) * WT_WARPS;
            constexpr int threads = WT_WARPS * 32;
            auto kern = cube ? stack_wtma_kernel<P, true> : stack_wtma_kernel<P, false>;
            AB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem));
            int per_sm = 0;
            AB_CUDA(ctx,
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
            if (per_sm < 1) per_sm = 1;
            size_t grid = (size_t)ctx->sm_count * per_sm;
            const size_t need = (groups + WT_WARPS - 1) / WT_WARPS;
            if (grid > need) grid = need;
            kern<<<(unsigned)grid, threads, smem, stream>>>(ft, tm, groups, sl, sh, max_iter, d_out,
                                                            d_rej);
            AB_LAUNCH_CHECK(ctx);
        }
        const size_t done = groups * 32;
        if (done < npix) {  // ragged tail (or everything, for unaligned frames): plain loads
            FrameTable tail;
            for (int f = 0; f < AB_MAX_FRAMES; ++f) tail.p[f] = f < P ? ft.p[f] + done : nullptr;
            return launch_ldg<P>(ctx, tail, npix - done, sl, sh, max_iter, d_out + done, d_rej,
                                 stream);
        }
        return AB_OK;
    }
    bool aligned = true;
    for (int f = 0; f < P; ++f)
        aligned = aligned && ((reinterpret_cast<uintptr_t>(ft.p[f]) & 15) == 0);
    const size_t full_tiles = aligned && ctx->stack_variant == 3 ? npix / TILE : 0;
    if (full_tiles) {
        // constant frame stride (a [P, rows, cols] cube) -> one 2-D tensor-map TMA per tile
        bool cube = P > 1 && !ctx->force_tma1d;
        ptrdiff_t stride = P > 1 ? ft.p[1] - ft.p[0] : 0;
        for (int f = 1; cube && f < P; ++f) cube = (ft.p[f] - ft.p[f - 1]) == stride;
        cube = cube && stride > 0 && (stride % 4) == 0 && (size_t)stride >= npix &&
               full_tiles * TILE < (size_t)1 << 31;
        TmapParam tm{};
        if (cube) cube = ab_make_stack_tmap(ctx, ft.p[0], npix, (size_t)stride, P, TILE, &tm);
        const size_t smem = sizeof(StackTmaSmem<P>);
        constexpr int threads = A_THREADS + B_THREADS + PROD_THREADS;
        auto kern = cube ? stack_tma_kernel<P, true> : stack_tma_kernel<P, false>;
        AB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
        int per_sm = 0;
        AB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
        if (per_sm < 1) per_sm = 1;
        size_t grid = (size_t)ctx->sm_count * per_sm;
        if (grid > full_tiles) grid = full_tiles;
        kern<<<(unsigned)grid, threads, smem, stream>>>(ft, tm, full_tiles, sl, sh, max_iter, d_out,
                                                        d_rej, ctx->stack_dbg);
        AB_LAUNCH_CHECK(ctx);
    }
    const size_t done = full_tiles * TILE;
    if (done < npix) {  // ragged tail (or everything, for unaligned frames): LDG kernel
        FrameTable tail;
        for (int f = 0; f < AB_MAX_FRAMES; ++f) tail.p[f] = f < P ? ft.p[f] + done : nullptr;
        return launch_ldg<P>(ctx, tail, npix - done, sl, sh, max_iter, d_out + done, d_rej, stream);
    }
    return AB_OK;
}

}  // namespace abk
