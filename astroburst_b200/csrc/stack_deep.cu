// stack_deep.cu -- sigma-clipped stack for frame counts without a register kernel (33..256 frames,
// BASELINE config 5: 256 x 4096^2).
//
// Reference semantics: stack_images / sigma_clip_combine,
// src-tauri/src/core/stacking/combine.rs:14-92,160-182 (align=false core loop).
//
// One thread = one output pixel; the N samples of a pixel are processed as K = ceil(N/32) chunks.
//
//  phase 1  per chunk, register-resident: 32 coalesced streaming loads (the next chunk's loads are
//    in flight while this one is sorted), non-finite -> +inf, the 185-comparator sorting network of
//    the register kernels (stack_kernels.cuh), then the sorted chunk is written to the thread's
//    column in shared memory (element e of lane l at word e*32 + l: every dynamic access below is
//    bank-conflict free) and the f64 sums S1 = sum (v - c), S2 = sum (v - c)^2 about a fixed centre
//    c (the first chunk's median) are accumulated.
//  phase 2  iteration 0 (combine.rs:37-48) on the K sorted lists: the upper median is the element
//    of rank n/2 of their union -- lower-bound binary searches for a pivot near it (mean of the
//    chunk medians), then a K-way merge walk over the few ranks in between; the upper-median MAD
//    is the rank n/2 element of the 2K deviation sequences that fan out left and right of the
//    median, found the same way.
//  phase 3  the clip (combine.rs:65-82): dev = v - center is monotone in v, so every list loses a
//    prefix and a suffix; removed samples are subtracted from S1, S2.  Iterations >= 1
//    (combine.rs:49-61) need no pass over the data: mean = c + S1/len exactly (the f64 sum of the
//    reference is exact for f32 samples of comparable magnitude), and the sample variance follows
//    from S2 - 2 d S1 + len d^2 with d = mean - c, which agrees with the reference's sequential
//    f64 accumulation to ~1e-15 relative before the cast to f32.
#include "stack_kernels.cuh"

namespace {

using abk::CeConst;

constexpr int CH = 32;           // samples per chunk
constexpr int MAX_CHUNKS = AB_MAX_FRAMES / CH;
constexpr int DEEP_MN = 2, DEEP_MD = 3;  // IMAD share of the compare-exchanges (tuned at r1: 4.05 ms)
// warps per block.  At K = 8 (225..256 frames) a warp's columns take 32 KiB: three 2-warp blocks fit
// an SM (each block also reserves 1 KiB), ONE 7-warp block fits seven warps -- occupancy is what
// bounds this kernel, so that layout is used there.
template <int K>
constexpr int deep_warps() {
    return K == 8 ? 7 : 2;
}

// 8 small integers (0..255) packed in one 64-bit register: per-list counters without local memory
struct Pack8 {
    unsigned long long v;
    __device__ __forceinline__ int get(int k) const { return (int)((v >> (8 * k)) & 0xffull); }
    __device__ __forceinline__ void set(int k, int x) {
        v = (v & ~(0xffull << (8 * k))) | ((unsigned long long)(x & 0xff) << (8 * k));
    }
    __device__ __forceinline__ void add(int k, int d) { v += (unsigned long long)(long long)d << (8 * k); }
};

// element i of list k in this lane's column
__device__ __forceinline__ float at(const float *col, int k, int i) { return col[(k * CH + i) * 32]; }

// Lock-step branchless searches: cnt[k] <- number of leading elements of list k's range that
// satisfy the (monotone) predicate.  Six fixed steps; the K lists are independent, so their
// shared-memory loads overlap.
//   pred(k, idx) must be true for a prefix of [base[k], limit[k]) and false after it.
template <int K, typename Pred>
__device__ __forceinline__ void prefix_counts(const int (&base)[K], const int (&limit)[K],
                                              int (&cnt)[K], Pred pred) {
#pragma unroll
    for (int k = 0; k < K; ++k) cnt[k] = 0;
#pragma unroll
    for (int step = CH; step >= 1; step >>= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int idx = base[k] + cnt[k] + step - 1;
            if (idx < limit[k] && pred(k, idx)) cnt[k] += step;
        }
    }
}

// Value of rank r (0-based) of the union of the K sorted lists (n[k] elements each): counts of
// elements below a pivot near the target, then a K-way merge walk over the ranks in between.
template <int K>
__device__ __forceinline__ float select_rank(const float *col, const int (&n)[K], int n_tot, int r,
                                             float pivot) {
    int zero[K], c[K];
#pragma unroll
    for (int k = 0; k < K; ++k) zero[k] = 0;
    auto below = [&](float p) {
        prefix_counts<K>(zero, n, c, [&](int k, int idx) { return at(col, k, idx) < p; });
        int t = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) t += c[k];
        return t;
    };
    int c_less = below(pivot);
    // one secant-like refinement: move the pivot to the element the count is off by, inside the
    // list whose own count is largest (any value works as a pivot; only the walk length changes)
    {
        const int off = r - c_less;  // > 0: pivot too low
        if (off > 2 || off < -2) {
            int kb = 0;
#pragma unroll
            for (int k = 1; k < K; ++k)
                if (n[k] > n[kb]) kb = k;
            int ck = 0, nk = 0;
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (k == kb) {
                    ck = c[k];
                    nk = n[k];
                }
            // the list holds ~nk/n_tot of the samples: shift its index by the same share of `off`
            int idx = ck + (off * nk + (off > 0 ? n_tot - 1 : -(n_tot - 1))) / n_tot;
            idx = max(0, min(nk - 1, idx));
            pivot = at(col, kb, idx);
            c_less = below(pivot);
        }
    }
    Pack8 pos{0};
#pragma unroll
    for (int k = 0; k < K; ++k) pos.set(k, c[k]);
    if (c_less <= r) {  // walk forward over the elements >= pivot
        int steps = r - c_less;
#pragma unroll 1
        for (;;) {
            float best = CUDART_INF_F;
            int arg = -1;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int i = pos.get(k);
                const float v = i < n[k] ? at(col, k, i) : CUDART_INF_F;
                if (i < n[k] && (arg < 0 || v < best)) {
                    best = v;
                    arg = k;
                }
            }
            if (steps == 0 || arg < 0) return best;
            pos.add(arg, 1);
            --steps;
        }
    }
    int steps = c_less - 1 - r;  // walk backward over the elements < pivot
#pragma unroll 1
    for (;;) {
        float best = -CUDART_INF_F;
        int arg = -1;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int i = pos.get(k);
            const float v = i > 0 ? at(col, k, i - 1) : -CUDART_INF_F;
            if (i > 0 && (arg < 0 || v > best)) {
                best = v;
                arg = k;
            }
        }
        if (steps == 0 || arg < 0) return best;
        pos.add(arg, -1);
        --steps;
    }
}

// Rank r2 element of the deviations |v - med| (f32, as combine.rs:41-44).  Per list the samples
// >= med give an ascending deviation sequence to the right of q = lower_bound(med), the samples
// < med one to the left; [lb, ub) is the window of samples whose deviation is < the pivot.
template <int K>
__device__ __forceinline__ float select_mad(const float *col, const int (&n)[K], float med, int r2,
                                            float pivot) {
    int zero[K], q[K], cr[K], cl[K];
#pragma unroll
    for (int k = 0; k < K; ++k) zero[k] = 0;
    prefix_counts<K>(zero, n, q, [&](int k, int idx) { return at(col, k, idx) < med; });
    auto inside = [&](float p) {
        // right: leading samples of [q, n) with (v - med) < p; left: leading samples of [0, q)
        // with (med - v) >= p (the deviation decreases with the index there)
        prefix_counts<K>(q, n, cr, [&](int k, int idx) { return at(col, k, idx) - med < p; });
        prefix_counts<K>(zero, q, cl, [&](int k, int idx) { return !(med - at(col, k, idx) < p); });
        int t = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) t += cr[k] + (q[k] - cl[k]);
        return t;
    };
    int cnt = inside(pivot);
    if ((cnt - r2 > 3 || r2 - cnt > 3) && cnt > 0 && pivot > 0.0f) {
        // deviations are roughly uniform near the MAD: rescale the pivot to the wanted count
        pivot = pivot * ((float)r2 + 0.5f) / (float)cnt;
        cnt = inside(pivot);
    }
    Pack8 lb{0}, ub{0};
#pragma unroll
    for (int k = 0; k < K; ++k) {
        lb.set(k, cl[k]);
        ub.set(k, q[k] + cr[k]);
    }
    if (cnt <= r2) {  // grow the windows: pop the smallest outside deviation
        int steps = r2 - cnt;
#pragma unroll 1
        for (;;) {
            float best = CUDART_INF_F;
            int arg = -1;  // k for a right head, k + 8 for a left head
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int u = ub.get(k), l = lb.get(k);
                const float dr = u < n[k] ? at(col, k, u) - med : CUDART_INF_F;
                if (u < n[k] && (arg < 0 || dr < best)) {
                    best = dr;
                    arg = k;
                }
                const float dl = l > 0 ? med - at(col, k, l - 1) : CUDART_INF_F;
                if (l > 0 && (arg < 0 || dl < best)) {
                    best = dl;
                    arg = k + 8;
                }
            }
            if (steps == 0 || arg < 0) return best;
            if (arg < 8) ub.add(arg, 1);
            else lb.add(arg - 8, -1);
            --steps;
        }
    }
    int steps = cnt - 1 - r2;  // shrink the windows: pop the largest inside deviation
#pragma unroll 1
    for (;;) {
        float best = -CUDART_INF_F;
        int arg = -1;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int u = ub.get(k), l = lb.get(k);
            const float dr = u > q[k] ? at(col, k, u - 1) - med : -CUDART_INF_F;
            if (u > q[k] && (arg < 0 || dr > best)) {
                best = dr;
                arg = k;
            }
            const float dl = l < q[k] ? med - at(col, k, l) : -CUDART_INF_F;
            if (l < q[k] && (arg < 0 || dl > best)) {
                best = dl;
                arg = k + 8;
            }
        }
        if (steps == 0 || arg < 0) return best;
        if (arg < 8) ub.add(arg, -1);
        else lb.add(arg - 8, 1);
        --steps;
    }
}

// PRE = false is the plain stack (no calibration terms, no median-only mode compiled in).
template <int K, bool PRE>
__global__ void __launch_bounds__(deep_warps<K>() * 32)
stack_deep_kernel(const __grid_constant__ FrameTable ft, int n_frames, size_t npix,
                  float sigma_low, float sigma_high, unsigned max_iter, CeConst kc,
                  const __grid_constant__ DeepPre pre, float *__restrict__ out,
                  unsigned long long *__restrict__ rejected_total) {
    extern __shared__ __align__(16) float deep_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float *col = deep_smem + warp * (K * CH * 32) + lane;
    unsigned local_rej = 0;
    const bool has_bias = PRE && pre.bias != nullptr;
    const bool has_dark = PRE && pre.dark != nullptr;
    const bool median_only = PRE && pre.median_only != 0;

    const size_t n_groups = (npix + 31) / 32;
    constexpr int DEEP_WARPS = deep_warps<K>();
    const size_t warp_global = (size_t)blockIdx.x * DEEP_WARPS + warp;
    const size_t warp_stride = (size_t)gridDim.x * DEEP_WARPS;
    for (size_t g = warp_global; g < n_groups; g += warp_stride) {
        const size_t p = g * 32 + lane;
        const bool active = p < npix;

        // ---------------------------------------------------------------- phase 1
        Pack8 npk{0};  // (the chunk loop stays rolled: one copy of the sorting network)
        int n_tot = 0;
        bool have_c = false;
        double cd = 0.0, S1 = 0.0, S2 = 0.0;
        float pivot_acc = 0.0f, spread_acc = 0.0f;
        int pivot_cnt = 0;
        // calibration.rs:160-162,199-206: frame - bias, then - dark * ratio (median-combine callers)
        const float pre_b = (active && has_bias) ? pre.bias[p] : 0.0f;
        const float pre_d = (active && has_dark) ? pre.dark[p] * pre.dark_ratio : 0.0f;
        auto fetch = [&](int f) -> float {
            if (!(active && f < n_frames)) return CUDART_INF_F;
            float v = __ldcs(ft.p[f] + p);
            if (has_bias) v -= pre_b;
            if (has_dark) v -= pre_d;
            return v;
        };
        float nxt[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) nxt[j] = fetch(j);
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            float s[CH];
            int nk = 0;
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const bool fin = fabsf(nxt[j]) < CUDART_INF_F;
                nk += fin ? 1 : 0;
                s[j] = fin ? nxt[j] : CUDART_INF_F;
            }
            if (k + 1 < K) {  // next chunk's loads fly while this one is sorted
#pragma unroll
                for (int j = 0; j < CH; ++j) nxt[j] = fetch((k + 1) * CH + j);
            }
            abk::sort_network<CH, DEEP_MN, DEEP_MD>(s, kc);
            float *dst = col + (size_t)k * CH * 32;
#pragma unroll
            for (int j = 0; j < CH; ++j) dst[j * 32] = s[j];
            npk.set(k, nk);
            n_tot += nk;
            if (nk > 0) {
                const float mk = dst[(nk >> 1) * 32];  // chunk median (own column: no sync needed)
                pivot_acc += mk;
                spread_acc += dst[((3 * nk) >> 2) * 32] - dst[(nk >> 2) * 32];
                ++pivot_cnt;
                if (!have_c) {
                    have_c = true;
                    cd = (double)mk;
                }
            }
            if (!median_only) {
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    if (j < nk) {
                        const double d = (double)s[j] - cd;
                        S1 += d;
                        S2 += d * d;
                    }
                }
            }
        }

        int n[K];
#pragma unroll
        for (int k = 0; k < K; ++k) n[k] = npk.get(k);

        float result = 0.0f;
        unsigned rejected = 0;
        if (median_only) {  // median_combine_row_major, calibration.rs:108-116
            if (n_tot > 0)
                result = select_rank<K>(col, n, n_tot, n_tot >> 1, pivot_acc / (float)pivot_cnt);
        } else if (n_tot == 1) {  // combine.rs:24-26
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (n[k]) result = at(col, k, 0);
        } else if (n_tot >= 2) {
            int len = n_tot;
            float last_center = 0.0f;
            int a[K], b[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                a[k] = 0;
                b[k] = n[k];
            }
            for (unsigned it = 0; it < max_iter; ++it) {
                if (len < 2) break;
                float center, sigma;
                if (it == 0) {
                    // ------------------------------------------------------------ phase 2
                    const float inv = 1.0f / (float)pivot_cnt;
                    const float med = select_rank<K>(col, n, n_tot, n_tot >> 1, pivot_acc * inv);
                    const float mad =
                        select_mad<K>(col, n, med, n_tot >> 1, 0.5f * spread_acc * inv);
                    center = med;
                    sigma = fmaxf((float)((double)mad * abk::MAD_TO_SIGMA), 1e-10f);
                } else {
                    const double mean = abk::div_small((double)len * cd + S1, len);
                    const double dm = mean - cd;
                    double ss = (S2 - 2.0 * dm * S1) + (double)len * dm * dm;
                    ss = fmax(ss, 0.0);
                    center = (float)mean;
                    sigma = fmaxf((float)sqrt(abk::div_small(ss, len > 2 ? len - 1 : 1)), 1e-10f);
                }
                last_center = center;
                // ---------------------------------------------------------------- phase 3
                const float lo = -sigma_low * sigma;
                const float hi = sigma_high * sigma;
                int removed = 0;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    int ak = a[k], bk = b[k];
#pragma unroll 1
                    while (ak < bk) {
                        const float v = at(col, k, ak);
                        if (v - center >= lo) break;
                        const double d = (double)v - cd;
                        S1 -= d;
                        S2 -= d * d;
                        ++ak;
                        ++removed;
                    }
#pragma unroll 1
                    while (bk > ak) {
                        const float v = at(col, k, bk - 1);
                        if (v - center <= hi) break;
                        const double d = (double)v - cd;
                        S1 -= d;
                        S2 -= d * d;
                        --bk;
                        ++removed;
                    }
                    a[k] = ak;
                    b[k] = bk;
                }
                rejected += (unsigned)removed;
                len -= removed;
                if (removed == 0) break;
            }
            if (len == 0)
                result = isfinite(last_center) ? last_center : 0.0f;  // combine.rs:85-88
            else
                result = (float)abk::div_small((double)len * cd + S1, len);
        }
        if (active) {
            __stcs(out + p, result);
            local_rej += rejected;
        }
    }
    const unsigned w = __reduce_add_sync(0xffffffffu, local_rej);
    if (lane == 0 && w) atomicAdd(rejected_total, (unsigned long long)w);
}

}  // namespace

template <int K, bool PRE>
static ab_status launch_deep_k(ab_ctx *ctx, const FrameTable &ft, int n_frames, size_t npix, float sl,
                               float sh, unsigned max_iter, const DeepPre &pre, float *d_out,
                               unsigned long long *d_rej, cudaStream_t stream) {
    constexpr int DEEP_WARPS = deep_warps<K>();
    const size_t smem = (size_t)DEEP_WARPS * K * CH * 32 * sizeof(float);
    auto kern = stack_deep_kernel<K, PRE>;
    AB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    AB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DEEP_WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    const size_t need = (npix + DEEP_WARPS * 32 - 1) / (DEEP_WARPS * 32);
    size_t grid = (size_t)ctx->sm_count * per_sm;  // persistent: one wave, groups strided
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, DEEP_WARPS * 32, smem, stream>>>(ft, n_frames, npix, sl, sh, max_iter,
                                                           CeConst{1, -1}, pre, d_out, d_rej);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

ab_status ab_stack_launch_deep(ab_ctx *ctx, const FrameTable &ft, int n_frames, size_t npix,
                               float sl, float sh, unsigned max_iter, const DeepPre &pre,
                               float *d_out, unsigned long long *d_rej, cudaStream_t stream) {
    const int K = (n_frames + CH - 1) / CH;
    const bool use_pre = pre.bias || pre.dark || pre.median_only;
#define AB_DEEP_CASE(k)                                                                         \
    case k:                                                                                     \
        return use_pre ? launch_deep_k<k, true>(ctx, ft, n_frames, npix, sl, sh, max_iter, pre,     \
                                                d_out, d_rej, stream)                           \
                       : launch_deep_k<k, false>(ctx, ft, n_frames, npix, sl, sh, max_iter, pre,    \
                                                 d_out, d_rej, stream);
    switch (K) {
        AB_DEEP_CASE(1) AB_DEEP_CASE(2) AB_DEEP_CASE(3) AB_DEEP_CASE(4) AB_DEEP_CASE(5)
        AB_DEEP_CASE(6) AB_DEEP_CASE(7) AB_DEEP_CASE(8)
        default:
            break;
    }
#undef AB_DEEP_CASE
    return ab_fail(ctx, AB_ERR_UNSUPPORTED, "deep stack: %d frames", n_frames);
}
