// stack.cu -- sigma-clipped frame stacking for sm_100a.
//
// Reference semantics: stack_images / sigma_clip_combine,
// src-tauri/src/core/stacking/combine.rs:14-92,160-182 (align=false core loop).
//
// Design (HBM-bound integer/compare work, no GEMM shape):
//  * one thread owns one output pixel; its <=P samples live in registers, loaded with one
//    fully coalesced streaming load per frame plane (32 lanes x 4 B = one 128 B line);
//  * the samples are sorted once with a Batcher odd-even merge network (191 compare-exchanges
//    for P=32, branch-free FMNMX pairs).  Non-finite samples are mapped to +inf so they sort
//    last and the finite count n is the logical length (combine.rs:170-175);
//  * iteration 0 (upper median + upper-median MAD, combine.rs:37-48) falls out of the sorted
//    order: med = s[n/2]; |v-med| is V-shaped over the sorted samples, so the MAD is the
//    (n/2+1)-th smallest of two sorted runs = min_t max(med - s[t], s[t+n/2] - med);
//  * survivors of every clip are a contiguous run of the sorted samples (dev = v - center is
//    monotone in v), tracked as a bit mask; mean / sample variance are f64 sums over the mask
//    in ascending order (combine.rs:50-60 sums in select_nth-permuted order: identical
//    whenever the f64 sum of f32 values is exact, i.e. for any dynamic range < 2^29);
//  * all f32 compares use the reference's exact expressions (dev >= -sl*sigma && dev <= sh*sigma,
//    sigma floors 1e-10, f64 -> f32 casts), compiled with -fmad=false.
//
// A shared-memory variant (any n <= 256, dynamic indices) covers deep stacks.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <utility>

#include <math_constants.h>

#include "ab_internal.h"

namespace {

constexpr double MAD_TO_SIGMA = 1.4826;  // types/constants.rs:7
constexpr int STACK_THREADS = 256;

__device__ __forceinline__ void ce(float &a, float &b) {
    const float lo = fminf(a, b);
    const float hi = fmaxf(a, b);
    a = lo;
    b = hi;
}

// Batcher odd-even merge sort network for any compile-time P, generated as a constexpr list of
// compare-exchange pairs and expanded with a fold expression, so every index is a literal and
// the array provably stays in registers (loop unrolling alone left rolled loops with
// select-chain "dynamic register indexing" in the SASS).
struct CePair {
    short a, b;
};
template <int P>
struct OemNet {
    static constexpr int kMax = P * 12;
    CePair pr[kMax];
    int count;
    constexpr OemNet() : pr{}, count(0) {
        for (int p = 1; p < P; p <<= 1)
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
__device__ __forceinline__ void sort_network_impl(float (&s)[P],
                                                  std::integer_sequence<int, I...>) {
    (ce(s[OemHolder<P>::net.pr[I].a], s[OemHolder<P>::net.pr[I].b]), ...);
}
template <int P>
__device__ __forceinline__ void sort_network(float (&s)[P]) {
    sort_network_impl<P>(s, std::make_integer_sequence<int, OemHolder<P>::net.count>{});
}

// Sorts a bitonic (here: V-shaped) register sequence ascending; K is the compile-time stride.
template <int P, int K>
__device__ __forceinline__ void bitonic_merge_step(float (&d)[P]) {
#pragma unroll
    for (int i = 0; i < P; ++i)
        if ((i & K) == 0) ce(d[i], d[i + K]);
    if constexpr (K > 1) bitonic_merge_step<P, K / 2>(d);
}
template <int P>
__device__ __forceinline__ void bitonic_merge(float (&d)[P]) {
    bitonic_merge_step<P, P / 2>(d);
}

// s[idx] for a runtime idx, register array kept in registers: binary select tree.  The selects
// are PTX selp so NVVM does not rewrite the chain into an indexed local-memory load.
__device__ __forceinline__ float selp(float a, float b, bool take_b) {
    float r;
    asm("{ .reg .pred p; setp.ne.b32 p, %3, 0; selp.f32 %0, %2, %1, p; }"
        : "=f"(r)
        : "f"(a), "f"(b), "r"((int)take_b));
    return r;
}
template <int P>
__device__ __forceinline__ float select_reg(const float (&s)[P], int idx) {
    if constexpr (P == 1) {
        return s[0];
    } else {
        float t[P / 2];
        const bool odd = idx & 1;
#pragma unroll
        for (int i = 0; i < P / 2; ++i) t[i] = selp(s[2 * i], s[2 * i + 1], odd);
        return select_reg<P / 2>(t, idx >> 1);
    }
}

template <int P>
struct MaskOf {
    using type = uint32_t;
};
template <>
struct MaskOf<64> {
    using type = uint64_t;
};

__device__ __forceinline__ int popc_mask(uint32_t m) { return __popc(m); }
__device__ __forceinline__ int popc_mask(uint64_t m) { return __popcll(m); }

// Post-sort part of sigma_clip_combine on a register-resident ascending array whose entries
// j >= n are +inf.  Returns the combined value; `rejected` gets the per-pixel reject count.
template <int P>
__device__ __forceinline__ float clip_combine_sorted(float (&s)[P], int n, float sigma_low,
                                                     float sigma_high, unsigned max_iter,
                                                     unsigned &rejected) {
    using mask_t = typename MaskOf<P>::type;
    rejected = 0;
    if (n == 0) return 0.0f;        // combine.rs:21-23
    if (n == 1) return s[0];        // combine.rs:24-26

    const mask_t valid = (n >= P) ? ~mask_t(0) : ((mask_t(1) << n) - 1);
    mask_t alive = valid;
    int len = n;
    float last_center = nanf("");
    bool mean_known = false;
    double mean_cur = 0.0;

    // One clip step (combine.rs:65-82).  Entries j >= n are +inf and masked by `valid`.
    auto clip = [&](float center, float sigma) -> int {
        const float lo = -sigma_low * sigma;
        const float hi = sigma_high * sigma;
        mask_t keep = 0;
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const float dev = s[j] - center;
            if (dev >= lo && dev <= hi) keep |= (mask_t(1) << j);
        }
        keep &= alive;
        const int new_len = popc_mask(keep);
        const int removed = len - new_len;
        rejected += (unsigned)removed;
        alive = keep;
        len = new_len;
        return removed;
    };

    if (max_iter == 0) goto finish;
    {
        // iteration 0: upper median and upper-median MAD (combine.rs:37-48)
        float med, mad;
        if (n == P) {
            med = s[P / 2];
            mad = CUDART_INF_F;
#pragma unroll
            for (int t = 0; t < (P + 1) / 2; ++t)
                mad = fminf(mad, fmaxf(med - s[t], s[t + P / 2] - med));
            // (for even P the last window pairs s[P/2] with the absent s[P] = +inf: skipped)
        } else {
            const int m = n >> 1;
            med = select_reg<P>(s, m);
            float d[P];
#pragma unroll
            for (int j = 0; j < P; ++j) d[j] = fabsf(s[j] - med);  // V-shaped; +inf stays +inf
            bitonic_merge<P>(d);
            mad = select_reg<P>(d, m);
        }
        last_center = med;
        const float sigma0 = (float)fmax((double)mad * MAD_TO_SIGMA, 1e-10);
        if (clip(med, sigma0) == 0) goto finish;
    }
    for (unsigned it = 1; it < max_iter; ++it) {
        if (len < 2) break;
        double sum = 0.0;
#pragma unroll
        for (int j = 0; j < P; ++j)
            if ((alive >> j) & 1) sum += (double)s[j];
        const double nn = (double)len;
        const double mean = sum / nn;
        double ss = 0.0;
#pragma unroll
        for (int j = 0; j < P; ++j)
            if ((alive >> j) & 1) {
                const double dd = (double)s[j] - mean;
                ss += dd * dd;
            }
        const double variance = ss / fmax(nn - 1.0, 1.0);
        const float center = (float)mean;
        const float sigma = (float)fmax(sqrt(variance), 1e-10);
        last_center = center;
        if (clip(center, sigma) == 0) {
            mean_known = true;  // same survivor set, same summation order as this pass
            mean_cur = mean;
            break;
        }
    }

finish:
    if (len == 0) return isfinite(last_center) ? last_center : 0.0f;  // combine.rs:85-88
    if (mean_known) return (float)mean_cur;  // same set, same summation order as the last pass
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < P; ++j)
        if ((alive >> j) & 1) sum += (double)s[j];
    return (float)(sum / (double)len);
}

template <int P>
__global__ void __launch_bounds__(STACK_THREADS)
stack_reg_kernel(const __grid_constant__ FrameTable ft, int n_frames, size_t npix,
                 float sigma_low, float sigma_high, unsigned max_iter, float *__restrict__ out,
                 unsigned long long *__restrict__ rejected_total) {
    unsigned long long local_rej = 0;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix;
         p += (size_t)gridDim.x * blockDim.x) {
        float s[P];
        int n = 0;
#pragma unroll
        for (int f = 0; f < P; ++f) {
            float v = CUDART_INF_F;
            if (f < n_frames) {
                v = __ldcs(ft.p[f] + p);
                const bool fin = isfinite(v);
                n += fin ? 1 : 0;
                v = fin ? v : CUDART_INF_F;
            }
            s[f] = v;
        }
        sort_network<P>(s);
        unsigned rej;
        const float r = clip_combine_sorted<P>(s, n, sigma_low, sigma_high, max_iter, rej);
        __stcs(out + p, r);
        local_rej += rej;
    }
    // block-level reduction of the reject counter: one global atomic per block
    __shared__ unsigned long long sh_rej;
    if (threadIdx.x == 0) sh_rej = 0;
    __syncthreads();
    unsigned w = __reduce_add_sync(0xffffffffu, (unsigned)local_rej);
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(&sh_rej, (unsigned long long)w);
    __syncthreads();
    if (threadIdx.x == 0 && sh_rej) atomicAdd(rejected_total, sh_rej);
}

// ---------------------------------------------------------------------------------------------
// Deep-stack path (64 < n_frames <= 256): one thread per pixel, samples in a shared-memory
// column smem[i * BLOCK + tid] (conflict-free), bitonic sort with runtime loops, then the same
// post-sort algorithm with dynamic indices.  Correct for any n; a faster warp-cooperative
// version is future work.
constexpr int DEEP_THREADS = 64;

__device__ __forceinline__ float clip_combine_sorted_mem(const float *s, int stride, int n,
                                                         float sigma_low, float sigma_high,
                                                         unsigned max_iter, unsigned &rejected) {
    rejected = 0;
    if (n == 0) return 0.0f;
    if (n == 1) return s[0];
    int a = 0, b = n;  // alive run [a, b) of the ascending samples
    float last_center = nanf("");
    bool mean_known = false;
    double mean_cur = 0.0;
    for (unsigned it = 0; it < max_iter; ++it) {
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
            double sum = 0.0;
            for (int j = a; j < b; ++j) sum += (double)s[j * stride];
            const double nn = (double)len;
            const double mean = sum / nn;
            double ss = 0.0;
            for (int j = a; j < b; ++j) {
                const double dd = (double)s[j * stride] - mean;
                ss += dd * dd;
            }
            center = (float)mean;
            sigma = (float)fmax(sqrt(ss / fmax(nn - 1.0, 1.0)), 1e-10);
            mean_known = true;
            mean_cur = mean;
        }
        last_center = center;
        const float lo = -sigma_low * sigma;
        const float hi = sigma_high * sigma;
        // dev = v - center is monotone (non-decreasing) over the ascending samples and never NaN
        // for finite v, so the survivors are exactly what is left after peeling both ends; the
        // negated compares keep NaN bounds rejecting everything, like the reference's filter.
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
    if (len == 0) return isfinite(last_center) ? last_center : 0.0f;
    if (mean_known) return (float)mean_cur;
    double sum = 0.0;
    for (int j = a; j < b; ++j) sum += (double)s[j * stride];
    return (float)(sum / (double)len);
}

__global__ void __launch_bounds__(DEEP_THREADS)
stack_deep_kernel(const __grid_constant__ FrameTable ft, int n_frames, int n_pad, size_t npix,
                  float sigma_low, float sigma_high, unsigned max_iter, float *__restrict__ out,
                  unsigned long long *__restrict__ rejected_total) {
    extern __shared__ float sm[];
    float *col = sm + threadIdx.x;
    const int stride = DEEP_THREADS;
    unsigned long long local_rej = 0;
    for (size_t base = (size_t)blockIdx.x * blockDim.x; base < npix;
         base += (size_t)gridDim.x * blockDim.x) {
        const size_t p = base + threadIdx.x;
        const bool active = p < npix;
        int n = 0;
        for (int f = 0; f < n_pad; ++f) {
            float v = CUDART_INF_F;
            if (active && f < n_frames) {
                v = __ldcs(ft.p[f] + p);
                const bool fin = isfinite(v);
                n += fin ? 1 : 0;
                v = fin ? v : CUDART_INF_F;
            }
            col[f * stride] = v;
        }
        // bitonic sort, ascending, private column (no cross-thread hazards)
        for (int k = 2; k <= n_pad; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = 0; i < n_pad; ++i) {
                    const int l = i ^ j;
                    if (l > i) {
                        const float x = col[i * stride], y = col[l * stride];
                        const bool up = (i & k) == 0;
                        if (up ? (x > y) : (x < y)) {
                            col[i * stride] = y;
                            col[l * stride] = x;
                        }
                    }
                }
            }
        }
        if (active) {
            unsigned rej;
            const float r =
                clip_combine_sorted_mem(col, stride, n, sigma_low, sigma_high, max_iter, rej);
            __stcs(out + p, r);
            local_rej += rej;
        }
    }
    unsigned w = __reduce_add_sync(0xffffffffu, (unsigned)local_rej);
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(rejected_total, (unsigned long long)w);
}

template <int P>
ab_status launch_reg(ab_ctx *ctx, const FrameTable &ft, int n_frames, size_t npix, float sl,
                     float sh, unsigned max_iter, float *d_out, unsigned long long *d_rej,
                     cudaStream_t stream) {
    int per_sm = 0;
    AB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stack_reg_kernel<P>,
                                                               STACK_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    const size_t need = (npix + STACK_THREADS - 1) / STACK_THREADS;
    // whole waves: a multiple of SM count x resident CTAs, grid-stride over the rest
    size_t grid = (size_t)ctx->sm_count * per_sm * 4;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    stack_reg_kernel<P><<<(unsigned)grid, STACK_THREADS, 0, stream>>>(ft, n_frames, npix, sl, sh,
                                                                     max_iter, d_out, d_rej);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

}  // namespace

ab_status ab_stack_launch(ab_ctx *ctx, const FrameTable &ft, int n_frames, size_t npix,
                          float sigma_low, float sigma_high, unsigned max_iter, float *d_out,
                          unsigned long long *d_rejected, cudaStream_t stream) {
    AB_CUDA(ctx, cudaMemsetAsync(d_rejected, 0, sizeof(unsigned long long), stream));
    if (npix == 0) return AB_OK;
    if (n_frames <= 8)
        return launch_reg<8>(ctx, ft, n_frames, npix, sigma_low, sigma_high, max_iter, d_out,
                             d_rejected, stream);
    if (n_frames <= 16)
        return launch_reg<16>(ctx, ft, n_frames, npix, sigma_low, sigma_high, max_iter, d_out,
                              d_rejected, stream);
    if (n_frames <= 32)
        return launch_reg<32>(ctx, ft, n_frames, npix, sigma_low, sigma_high, max_iter, d_out,
                              d_rejected, stream);
    if (n_frames <= 64)
        return launch_reg<64>(ctx, ft, n_frames, npix, sigma_low, sigma_high, max_iter, d_out,
                              d_rejected, stream);
    if (n_frames > AB_MAX_FRAMES)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "n_frames %d > AB_MAX_FRAMES %d", n_frames,
                       AB_MAX_FRAMES);

    // deep path
    int n_pad = 128;
    while (n_pad < n_frames) n_pad <<= 1;
    const size_t smem = (size_t)n_pad * DEEP_THREADS * sizeof(float);
    AB_CUDA(ctx, cudaFuncSetAttribute(stack_deep_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    AB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stack_deep_kernel,
                                                               DEEP_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const size_t need = (npix + DEEP_THREADS - 1) / DEEP_THREADS;
    size_t grid = (size_t)ctx->sm_count * per_sm * 4;
    if (grid > need) grid = need;
    stack_deep_kernel<<<(unsigned)grid, DEEP_THREADS, smem, stream>>>(
        ft, n_frames, n_pad, npix, sigma_low, sigma_high, max_iter, d_out, d_rejected);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}
