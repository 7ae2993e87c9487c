// stack_kernels.cuh -- device code of the sigma-clipped frame stack (sm_100a).
//
// Reference semantics: stack_images / sigma_clip_combine,
// src-tauri/src/core/stacking/combine.rs:14-92,160-182 (align=false core loop).
//
// Kernel structure (one thread = one output pixel, P = exact frame count, samples in registers)
//
//  phase A  every pixel, straight-line, register-resident:
//    * P conflict-free LDS from a warp-private TMA-staged tile (stack_wtma_kernel) or P coalesced
//      streaming loads (stack_warp_kernel: unaligned frames, ragged tails);
//    * sorting network (185 compare-exchanges at P=32: two 60-comparator 16-sorters + Batcher's
//      odd-even merge); non-finite samples map to +inf and sort last (combine.rs:170-175 drops
//      them).  The network is ALU-pipe bound when every compare-exchange is an FMNMX pair
//      (64 lanes/clk/SM), so a fraction MN/MD of the compare-exchanges computes the upper output
//      on the FMA pipe instead: lo = FMNMX.min(a, b) and hi = a + b - lo on the integer bit
//      patterns (two IMADs; exact because lo is bit-identical to a or b).  Measured on B200
//      (tools/ubench_ce.cu): 5.65 SM-cycles per 32-sort with FMNMX pairs, 4.3 with a 1/2 mix;
//    * iteration 0 (combine.rs:37-48) from the sorted order: med = s[P/2] (upper median) and,
//      because |v-med| is V-shaped over sorted samples, the upper-median MAD is
//      min_t max(med - s[t], s[t+P/2] - med).  The same differences d[j] = s[j] - med feed the
//      exact f32 sums of the mean;
//    * the clip (combine.rs:65-82): dev = v - center is monotone in v, so nothing is rejected iff
//      d[0] >= -sigma_low*sigma and d[P-1] <= sigma_high*sigma.  ~80 % of pixels: the reference
//      loop stops (combine.rs:80-82) and the result is the f64 mean of all P samples.
//  phase B  every other pixel (something rejected at iteration 0, non-finite samples, or a sum
//    the f32 certificate cannot prove exact): the sorted samples are parked in a warp-private ring
//    in shared memory and finished 32 at a time, one per lane, with the dynamic-index version of
//    the algorithm (peel the rejects off both ends, mean / sample variance in f64 over the
//    surviving run, repeat).
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

// developer A/B knobs (python -m astroburst_b200.build --variant NAME KNOB=VALUE)
#ifndef AB_WQ_CAP
#define AB_WQ_CAP 44
#endif
#ifndef AB_DRAIN_AT
#define AB_DRAIN_AT 32
#endif
#ifndef AB_STATIC_PEEL
#define AB_STATIC_PEEL 1
#endif
#ifndef AB_STATIC_NUM
#define AB_STATIC_NUM 7
#define AB_STATIC_DEN 8
#endif
#ifndef AB_TMA_WARPS
#define AB_TMA_WARPS 20
#endif
#ifndef AB_LDG_BLOCKS
#define AB_LDG_BLOCKS 3
#endif
#ifndef AB_PROBE2
#define AB_PROBE2 1
#endif
#ifndef AB_PROXY_FENCE
#define AB_PROXY_FENCE 0
#endif

namespace abk {

constexpr double MAD_TO_SIGMA = 1.4826;  // types/constants.rs:7
// meta word of a parked pixel: a | b << 8 | it0 << 16 | flags
constexpr unsigned META_FINAL_ONLY = 0x80000000u;  // only the final mean of [0, b) is left
constexpr unsigned META_SIGMA0 = 0x40000000u;      // all P finite; iteration 0's sigma is in `aux`

// Runtime 1 and -1 (kernel parameters): the multiplier of the IMAD form must not be a
// compile-time constant, or ptxas folds the IMADs back into ALU-pipe IADD3s.
struct CeConst {
    int one, mone;
};

// Compare-exchange.  kImad = false: FMNMX pair (2 ALU-pipe instructions).  kImad = true: FMNMX
// min on the ALU pipe, hi = a + b - lo in integer arithmetic on the FMA pipe.  Operands are finite
// or +inf here (never NaN), so lo is always bit-identical to one of the inputs and the integer
// identity returns the other one exactly, for any signs.
template <bool kImad>
__device__ __forceinline__ void ce(float &a, float &b, const CeConst &k) {
    const float lo = fminf(a, b);
    if constexpr (kImad) {
        const int t = __float_as_int(a) * k.one + __float_as_int(b);  // independent of the FMNMX
        b = __int_as_float(__float_as_int(lo) * k.mone + t);
    } else {
        b = fmaxf(a, b);
    }
    a = lo;
}

// Sorting network for any compile-time P, generated as a constexpr list of compare-exchange pairs
// and expanded with a fold expression, so every index is a literal and the array provably stays
// in registers.
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
// compare-exchange I takes the IMAD form for an evenly spread fraction MN/MD of the network
template <int MN, int MD>
constexpr bool ce_is_imad(int i) {
    return MN > 0 && ((i * MN) % MD) + MN >= MD;
}
template <int P, int MN, int MD, int... I>
__device__ __forceinline__ void sort_network_impl(float (&s)[P], const CeConst &k,
                                                  std::integer_sequence<int, I...>) {
    (ce<ce_is_imad<MN, MD>(I)>(s[OemHolder<P>::net.pr[I].a], s[OemHolder<P>::net.pr[I].b], k), ...);
}
template <int P, int MN, int MD>
__device__ __forceinline__ void sort_network(float (&s)[P], const CeConst &k) {
    if constexpr (P > 1)
        sort_network_impl<P, MN, MD>(s, k,
                                     std::make_integer_sequence<int, OemHolder<P>::net.count>{});
}

// ---- x / n for a small positive integer n, correctly rounded, without the f64 division routine.
// r = RN(1/n) from a table; q = RN(x r); rem = x - n q exactly (fma); q' = RN(q + rem r) is the
// correctly rounded quotient (Markstein).  Checked against the hardware division for 5e8 operands
// over n = 1..256 on the host.
struct RcpTable {
    double r[AB_MAX_FRAMES + 1];
    constexpr RcpTable() : r{} {
        r[0] = 0.0;
        for (int i = 1; i <= AB_MAX_FRAMES; ++i) r[i] = 1.0 / (double)i;
    }
};
static __constant__ RcpTable g_rcp = RcpTable();
__device__ __forceinline__ double div_small(double x, int n) {
    const double r = g_rcp.r[n], dn = (double)n;
    const double q = x * r;
    return fma(fma(-q, dn, x), r, q);
}

// ---- f64 sums of the surviving run s[a..b) of the ascending samples.
//
// The reference accumulates `v as f64` (combine.rs:51,90).  Converting every sample costs one
// F2F.F64.F32 each, and that instruction issues at 16 lanes/clk/SM (measured), so the conversions --
// not the DADDs -- dominate the mean.  When the run is positive and spans at most a factor of two
// (s[b-1] <= 2 s[a], the normal case for sky-dominated pixels) every difference s[j] - c with c
// inside the run is exact in f32 (Sterbenz) and a multiple of ulp(s[a]); the one-signed partial
// sums stay exact in f32 as long as the total is below 2^24 ulps = 2 * 2^floor(log2 s[a]).  Then
//     sum = len * c + D        (exact in f64: < 2^31 ulps)
// is the exact real sum, which is also what the reference's f64 accumulation yields for such a run,
// so the result is bit-identical.  The certificate is checked per pixel; anything else takes the
// plain f64 accumulation (4 accumulators keyed by the index relative to a, combined pairwise).
__device__ __forceinline__ float next_pow2_above(float s0) {  // 2^(floor(log2 s0) + 1), s0 normal > 0
    return __uint_as_float((__float_as_uint(s0) & 0x7f800000u) + 0x00800000u);
}
__device__ __forceinline__ double sum_run_mem(const float *s, int stride, int a, int b) {
    const float s0 = s[a * stride];
    if (s0 >= 1.17549435e-38f && s[(b - 1) * stride] <= 2.0f * s0) {
        const int mid = (a + b) >> 1;
        const float c = s[mid * stride];
        float lo = 0.0f, hi = 0.0f;
#pragma unroll 1
        for (int j = a; j < mid; ++j) lo += s[j * stride] - c;
#pragma unroll 1
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
// center.  it0 == 0 runs the whole algorithm (combine.rs:14-92); with have_sigma0 iteration 0's
// center (last_center = upper median) and sigma (sigma0) come from phase A.
__device__ __forceinline__ float clip_combine_sorted_mem(const float *s, int stride, int n, int a,
                                                         int b, unsigned it0, float last_center,
                                                         bool have_sigma0, float sigma0,
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
            if (have_sigma0) {
                center = last_center;
                sigma = sigma0;
            } else {
                const int m = n >> 1;
                const float med = s[m * stride];
                float mad = CUDART_INF_F;
#pragma unroll 1
                for (int t = 0; t <= m; ++t) {
                    const float left = med - s[t * stride];
                    const float right = (t + m < n) ? s[(t + m) * stride] - med : CUDART_INF_F;
                    mad = fminf(mad, fmaxf(left, right));
                }
                center = med;
                sigma = fmaxf((float)((double)mad * MAD_TO_SIGMA), 1e-10f);
            }
        } else {
            const double mean = div_small(sum_run_mem(s, stride, a, b), len);
            const double ss = sumsq_run_mem(s, stride, a, b, mean);
            center = (float)mean;
            // (float)max(sd, 1e-10) == max((float)sd, (float)1e-10): rounding is monotone
            sigma = fmaxf((float)sqrt(div_small(ss, len > 2 ? len - 1 : 1)), 1e-10f);
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
    return (float)div_small(sum_run_mem(s, stride, a, b), len);
}

// Out-of-line: one copy of the dynamic-index code per kernel.
static __device__ __noinline__ float finish_dynamic(const float *s, int stride, int n, int a, int b,
                                                    unsigned it0, float last_center,
                                                    bool have_sigma0, float sigma0, unsigned &rej,
                                                    float sigma_low, float sigma_high,
                                                    unsigned max_iter) {
    return clip_combine_sorted_mem(s, stride, n, a, b, it0, last_center, have_sigma0, sigma0, rej,
                                   sigma_low, sigma_high, max_iter);
}

// Frame-uniform gaps (the NaN borders registration leaves: every pixel of a region misses the
// same frames) make n warp-uniform; up to MAX_MISSING dropped samples are then handled by
// register-resident code specialised on n instead of the dynamic-index code.
constexpr int MAX_MISSING = 8;
template <int P>
constexpr int static_missing() {
    return P >= 24 ? MAX_MISSING : 0;  // dropped samples phase B still handles in registers
}

// Phase B of a META_SIGMA0 pixel (n >= P - static_missing<P>() finite samples, something rejected
// at iteration 0).  The sorted samples are re-read into registers with 128-bit loads; the
// surviving run [a, b) almost always keeps the "core" [E, n-E), so the passes of iterations >= 1
// run over the core with literal register indices (no loop, address or predicate instructions)
// and only the few edge elements are visited with dynamic indices in shared memory.  Zones of
// the core, for every admissible n (median index n/2 in [ML, MH]):
//   A = [E, ML)       below the median: differences <= 0
//   M = [ML, MH]      around the median: differences split by sign
//   B = (MH, P-M-E)   above the median: differences >= 0
//   C = [P-M-E, P-E)  only when n == P (else a dynamic edge up to b)
// The core's exact f32 difference sums (see sum_run_mem) do not depend on the iteration and are
// taken once.  A run that loses core elements continues in the dynamic-index code at the same
// iteration.  col = the slot's row in the WarpQueue (stride 1, 16-byte aligned).
template <int P>
__device__ __forceinline__ float finish_static(const float *col, int n, float med, float sigma0,
                                               unsigned &rej, float sigma_low, float sigma_high,
                                               unsigned max_iter) {
    constexpr int M = static_missing<P>();
    constexpr int E = P >= 24 ? 3 : (P >= 12 ? 2 : (P >= 6 ? 1 : 0));
    constexpr int ML = (P - M) / 2, MH = P / 2;
    constexpr int C0 = E, CB = P - M - E, CC = P - E;
    static_assert(C0 <= ML && MH < CB && CB <= CC, "zone layout");
    const bool full = n == P;
    const int c1 = full ? CC : CB;  // the static core of this pixel is [C0, c1)
    float s[P];
#pragma unroll
    for (int j = 0; j + 4 <= P; j += 4) {
        const float4 v4 = *reinterpret_cast<const float4 *>(col + j);
        s[j] = v4.x;
        s[j + 1] = v4.y;
        s[j + 2] = v4.z;
        s[j + 3] = v4.w;
    }
#pragma unroll
    for (int j = P & ~3; j < P; ++j) s[j] = col[j];
    float lo_acc[2] = {0.0f, 0.0f}, hi_acc[2] = {0.0f, 0.0f};
#pragma unroll
    for (int j = C0; j < ML; ++j) lo_acc[j & 1] += s[j] - med;  // <= 0
#pragma unroll
    for (int j = ML; j <= MH; ++j) {
        const float d = s[j] - med;
        lo_acc[j & 1] += fminf(d, 0.0f);
        hi_acc[j & 1] += fmaxf(d, 0.0f);
    }
#pragma unroll
    for (int j = MH + 1; j < CB; ++j) hi_acc[j & 1] += s[j] - med;  // >= 0
    if (full) {
#pragma unroll
        for (int j = CB; j < CC; ++j) hi_acc[j & 1] += s[j] - med;
    }
    const float core_lo = lo_acc[0] + lo_acc[1], core_hi = hi_acc[0] + hi_acc[1];

    int a = 0, b = n;
    float center = med, sigma = sigma0;
    for (unsigned it = 0;;) {
        const float lo = -sigma_low * sigma;
        const float hi = sigma_high * sigma;
        // dev = v - center is monotone over the ascending samples: the rejects are a prefix and a
        // suffix of the run.  The first E candidates of either end are tested in registers
        // (a <= E and, for n == P, b >= P - E here); the loops only continue past them.
        int na = a, nb = b;
        if constexpr (E > 0 && AB_STATIC_PEEL) {
            int kl = 0, kr = 0;
#pragma unroll
            for (int j = 0; j < E; ++j) kl += !(s[j] - center >= lo) ? 1 : 0;
            na = max(a, kl);  // rejected prefix of [0, E) (monotone), never below a
            if (full) {
#pragma unroll
                for (int j = 0; j < E; ++j) kr += !(s[P - 1 - j] - center <= hi) ? 1 : 0;
                nb = min(b, P - kr);
            }
        }
#pragma unroll 1
        while (na < nb && (!AB_STATIC_PEEL || na >= E) && !(col[na] - center >= lo)) ++na;
#pragma unroll 1
        while (nb > na && (!AB_STATIC_PEEL || !full || nb <= P - E) &&
               !(col[nb - 1] - center <= hi))
            --nb;
        if (nb < na) nb = na;  // everything rejected from both ends
        const int removed = (b - a) - (nb - na);
        rej += (unsigned)removed;
        a = na;
        b = nb;
        const int len = b - a;
        // removed == 0 at iteration >= 1: the loop ends and the final mean is the mean just used
        // as center (same run); `center` holds its f32 value already
        if (removed == 0 && it > 0) return center;
        if (len == 0) return isfinite(center) ? center : 0.0f;  // combine.rs:85-88
        ++it;
        if (a > C0 || b < c1)  // the run lost core elements: rare, dynamic indices from here on
            return finish_dynamic(col, 1, n, a, b, removed ? it : 0xffffffffu, center, false, 0.0f,
                                  rej, sigma_low, sigma_high, max_iter);
        // mean of the run: the next iteration's center (combine.rs:49-53) or, when the loop is
        // over, the result (combine.rs:90-91).  Exact sum = core sums + edge differences when
        // the f32 certificate holds, else the f64 accumulation.
        float d_lo = core_lo, d_hi = core_hi;
#pragma unroll
        for (int j = 0; j < C0; ++j) d_lo += j >= a ? s[j] - med : 0.0f;
        if (full) {
#pragma unroll
            for (int j = CC; j < P; ++j) d_hi += j < b ? s[j] - med : 0.0f;
        } else {
#pragma unroll 1
            for (int j = c1; j < b; ++j) d_hi += col[j] - med;
        }
        const float sa = col[a];
        const float T = next_pow2_above(sa);
        double sum;
        if (sa >= 1.17549435e-38f && col[b - 1] <= 2.0f * sa && d_hi < T && -d_lo < T)
            sum = (double)len * (double)med + (double)(d_hi + d_lo);
        else
            sum = sum_run_mem(col, 1, a, b);
        const double mean = div_small(sum, len);
        center = (float)mean;
        if (removed == 0 || it >= max_iter || len < 2) return center;
        // sample standard deviation of the run (combine.rs:54-60)
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int j = C0; j < CB; ++j) {
            const double e = (double)s[j] - mean;
            acc[j & 3] += e * e;
        }
        if (full) {
#pragma unroll
            for (int j = CB; j < CC; ++j) {
                const double e = (double)s[j] - mean;
                acc[j & 3] += e * e;
            }
        }
#pragma unroll
        for (int j = 0; j < C0; ++j) {
            const double e = (double)s[j] - mean;
            acc[j & 3] += j >= a ? e * e : 0.0;
        }
        if (full) {
#pragma unroll
            for (int j = CC; j < P; ++j) {
                const double e = (double)s[j] - mean;
                acc[j & 3] += j < b ? e * e : 0.0;
            }
        } else {
#pragma unroll 1
            for (int j = c1; j < b; ++j) {
                const double e = (double)col[j] - mean;
                acc[1] += e * e;
            }
        }
        const double ss = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        sigma = fmaxf((float)sqrt(div_small(ss, len > 2 ? len - 1 : 1)), 1e-10f);
    }
}

// Iteration 0 and the no-reject mean on the first N sorted samples of s[] (all finite; s[N..P) are
// the +inf stand-ins of dropped samples).  Returns true when `result` is final (such a pixel
// rejected nothing); false when the pixel must be finished by phase B: result = last center,
// aux = iteration 0's sigma and meta = N << 8 | flags.
template <int P, int N>
__device__ __forceinline__ bool phase_a_tail(const float (&s)[P], float sigma_low, float sigma_high,
                                             unsigned max_iter, float &result, float &aux,
                                             unsigned &meta) {
    static_assert(N >= 2 && N <= P, "phase_a_tail needs 2 <= N <= P");
    // upper median, upper-median MAD (combine.rs:37-48); the differences d[j] = s[j] - med also
    // feed the exact f32 sums of the mean (see sum_run_mem)
    constexpr int H = N / 2;
    const float med = s[H];
    float lo_acc[2] = {0.0f, 0.0f}, hi_acc[2] = {0.0f, 0.0f};
    float mad[2] = {CUDART_INF_F, CUDART_INF_F};
    float d_first = 0.0f, d_last = 0.0f;
#pragma unroll
    for (int t = 0; t < (N + 1) / 2; ++t) {
        float left = 0.0f, right = 0.0f;
        if (t < H) {  // (t == H only for odd N: the window that starts at the median)
            const float dl = s[t] - med;  // <= 0
            lo_acc[t & 1] += dl;
            left = -dl;
            if (t == 0) d_first = dl;
        }
        if (t > 0) {  // (t == 0 pairs with the median itself: difference 0)
            const float dr = s[t + H] - med;  // >= 0
            hi_acc[t & 1] += dr;
            right = dr;
            if (t + H == N - 1) d_last = dr;
        }
        mad[t & 1] = fminf(mad[t & 1], fmaxf(left, right));
    }
    // (float)max(x, 1e-10) == max((float)x, (float)1e-10): rounding is monotone
    const float sigma = fmaxf((float)((double)fminf(mad[0], mad[1]) * MAD_TO_SIGMA), 1e-10f);
    const float lo = -sigma_low * sigma;
    const float hi = sigma_high * sigma;
    // nothing rejected (combine.rs:65-82) <=> the extreme samples pass; max_iter == 0 never clips
    const bool keep_all = max_iter == 0 || (d_first >= lo && d_last <= hi);
    const float d_lo = lo_acc[0] + lo_acc[1], d_hi = hi_acc[0] + hi_acc[1];
    const float T = next_pow2_above(s[0]);
    // (power-of-two N takes an all-f32 mean below: D / N must not underflow, hence the 1e-30 floor;
    // smaller samples go to phase B's f64 mean)
    constexpr float kMinS0 = (N & (N - 1)) == 0 ? 1e-30f : 1.17549435e-38f;
    const bool exact = s[0] >= kMinS0 && s[N - 1] <= 2.0f * s[0] && d_hi < T && -d_lo < T;
    if (keep_all && exact) {
        if constexpr ((N & (N - 1)) == 0) {
            // N = 2^k: (N med + D) / N = med + D / N with D / N exact in f32, so one f32 addition
            // rounds exactly like the f64 route's final cast -- no F2F / f64 instructions at all
            result = med + (d_hi + d_lo) * (1.0f / (float)N);
        } else {
            const double sum = (double)N * (double)med + (double)(d_hi + d_lo);
            result = (float)div_small(sum, N);
        }
        return true;
    }
    result = med;  // last_center carried into phase B
    aux = sigma;
    meta = ((unsigned)N << 8) | (keep_all ? META_FINAL_ONLY : META_SIGMA0);
    return false;
}

// Warp-uniform n < P: the tail specialised on n (see MAX_MISSING).
template <int P, int K>
__device__ __forceinline__ bool phase_a_uniform(const float (&s)[P], int n, float sigma_low,
                                                float sigma_high, unsigned max_iter, float &result,
                                                float &aux, unsigned &meta, bool &handled) {
    if constexpr (K > MAX_MISSING || P - K < 2) {
        handled = false;
        return false;
    } else {
        if (n == P - K) {
            handled = true;
            return phase_a_tail<P, P - K>(s, sigma_low, sigma_high, max_iter, result, aux, meta);
        }
        return phase_a_uniform<P, K + 1>(s, n, sigma_low, sigma_high, max_iter, result, aux, meta,
                                         handled);
    }
}

// Phase A on a register array of exactly P finite-or-+inf samples (n of them finite; n_uni = n
// when every lane of the warp has the same n, else -1).
template <int P, int MN, int MD>
__device__ __forceinline__ bool phase_a(float (&s)[P], int n, int n_uni, float sigma_low,
                                        float sigma_high, unsigned max_iter, const CeConst &k,
                                        float &result, float &aux, unsigned &meta) {
    aux = 0.0f;
    meta = 0u;
    if constexpr (P == 1) {
        result = n ? s[0] : 0.0f;  // combine.rs:21-26
        return true;
    } else {
        sort_network<P, MN, MD>(s, k);
        if (n == P) return phase_a_tail<P, P>(s, sigma_low, sigma_high, max_iter, result, aux, meta);
        if (n == n_uni) {  // warp-uniform branch
            bool handled;
            const bool done = phase_a_uniform<P, 1>(s, n, sigma_low, sigma_high, max_iter, result,
                                                    aux, meta, handled);
            if (handled) return done;
        }
        // ragged non-finite samples: rare, finished with dynamic indices (combine.rs:21-26 first)
        result = n ? s[0] : 0.0f;
        if (n <= 1) return true;
        meta = (unsigned)n << 8;
        return false;
    }
}

// ---------------------------------------------------------------------------------------------
// Warp-autonomous kernels: no block-wide synchronisation at all.
//
// Every warp walks its own sequence of 32-pixel groups.  Lanes whose pixel needs phase B park the
// sorted samples in a warp-private stack in shared memory (column = slot, conflict-free).  At the
// top of a group iteration -- where nothing but the loop state is live, so the register-resident
// phase B can be inlined without spills -- a warp whose stack holds >= DRAIN_AT (32) pixels
// finishes the top 32 of them at once, one per lane.  Phase B therefore runs every ~6th group with
// full lanes; while one warp is in its (FP64 / XU heavy) phase B the other warps of the SM keep
// the ALU and FMA pipes busy with their sorting networks: warps drift apart instead of marching in
// lock-step.  A group whose parked pixels do not fit the WQ_CAP - 31 or more free slots (rare:
// twice the average park rate) first empties the stack through the out-of-line dynamic code.
constexpr int WARPS_PER_BLOCK = 8;      // LDG kernel: 8 warps x AB_LDG_BLOCKS blocks per SM
constexpr int WQ_CAP = AB_WQ_CAP;       // slots per warp
constexpr int DRAIN_AT = AB_DRAIN_AT;   // inline drain threshold: a full warp of parked pixels
static_assert(DRAIN_AT <= 32 && DRAIN_AT <= WQ_CAP, "one parked pixel per lane");

// One row of sorted samples per slot.  ROW = 4 x odd words: a lane's row is 16-byte aligned and the
// rows of a quarter-warp start 4 banks apart, so parking and the register reload are conflict-free
// 128-bit accesses (P/4 instructions instead of P).
template <int P, int CAP = WQ_CAP>
struct WarpQueue {
    static constexpr int kCap = CAP;
    static constexpr int ROW = 4 * (2 * ((P + 7) / 8) + 1);
    alignas(16) float v[CAP][ROW];
    // per slot: {last center, iteration-0 sigma, meta word, pixel index relative to the launch
    // (npix < 2^32 per launch)} as one 16-byte record: one STS.128 / LDS.128
    alignas(16) uint4 info[CAP];
};

// finish the `cnt` (<= 32) parked pixels of slots [first, first + cnt); lane i takes slot first + i.  kStatic: register-resident
// phase B for META_SIGMA0 pixels (inlined once per kernel); otherwise everything goes through the
// shared out-of-line dynamic-index code.
template <int P, bool kStatic>
__device__ __forceinline__ void drain_impl(WarpQueue<P> &q, unsigned first, unsigned cnt, int lane,
                                           float sigma_low, float sigma_high, unsigned max_iter,
                                           float *__restrict__ out, unsigned &local_rej) {
    if ((unsigned)lane < cnt) {
        const unsigned slot = first + lane;
        const uint4 info = q.info[slot];
        const float lastc = __uint_as_float(info.x), aux = __uint_as_float(info.y);
        const unsigned meta = info.z;
        const float *col = &q.v[slot][0];
        unsigned rej = 0;
        float r;
        const int n = (int)((meta >> 8) & 0x1ffu);
        if (kStatic && (meta & META_SIGMA0) && n >= P - static_missing<P>()) {
            r = finish_static<P>(col, n, lastc, aux, rej, sigma_low, sigma_high, max_iter);
        } else {  // ragged non-finite samples, or only the f64 mean of everything is left
            const unsigned it0 = (meta & META_FINAL_ONLY) ? 0xffffffffu : 0u;  // skips the clip loop
            r = finish_dynamic(col, 1, n, 0, n, it0, lastc, (meta & META_SIGMA0) != 0, aux, rej,
                               sigma_low, sigma_high, max_iter);
        }
        __stcs(out + info.w, r);
        local_rej += rej;
    }
    __syncwarp();
}
template <int P>
__device__ __noinline__ void drain_dynamic(WarpQueue<P> &q, unsigned cnt, int lane, float sigma_low,
                                           float sigma_high, unsigned max_iter,
                                           float *__restrict__ out, unsigned &local_rej) {
    for (unsigned first = 0; first < cnt; first += 32u)  // (cnt <= WQ_CAP: at most two rounds)
        drain_impl<P, false>(q, first, cnt - first < 32u ? cnt - first : 32u, lane, sigma_low,
                             sigma_high, max_iter, out, local_rej);
}

// Non-finite samples -> +inf (they sort last and are not counted in n).  probe is NaN iff any
// sample of the pixel is non-finite (x * 0 accumulates NaN for +-inf and NaN).
template <int P>
__device__ __forceinline__ int drop_non_finite(float (&s)[P], float probe) {
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
    return n;
}

// Shared tail of both kernels: phase A on the registers, then store or park.  Every lane runs
// phase A (inactive lanes of a ragged last group carry zeros); n_uni is found with two votes only
// when some lane of the warp dropped samples.
template <int P, int MN, int MD>
__device__ __forceinline__ void process_group(float (&s)[P], int n, bool active, unsigned p, int lane,
                                              WarpQueue<P> &q, unsigned &count, unsigned &local_rej,
                                              float sigma_low, float sigma_high, unsigned max_iter,
                                              const CeConst &k, float *__restrict__ out) {
    int n_uni = P;
    if (__any_sync(0xffffffffu, n != P)) {
        const int n0 = __shfl_sync(0xffffffffu, n, 0);
        n_uni = __all_sync(0xffffffffu, n == n0) ? n0 : -1;
    }
    float result, aux;
    unsigned meta;
    const bool done =
        phase_a<P, MN, MD>(s, n, n_uni, sigma_low, sigma_high, max_iter, k, result, aux, meta);
    if (active && done) __stcs(out + p, result);
    const bool park = active && !done;
    const unsigned pmask = __ballot_sync(0xffffffffu, park);
    if (pmask) {
        const unsigned c = __popc(pmask);
        if (count + c > (unsigned)WQ_CAP) {  // rare: make room first
            drain_dynamic<P>(q, count, lane, sigma_low, sigma_high, max_iter, out, local_rej);
            count = 0;
        }
        if (park) {
            const unsigned slot = count + __popc(pmask & ((1u << lane) - 1u));
            float *row = &q.v[slot][0];
#pragma unroll
            for (int f = 0; f + 4 <= P; f += 4)
                *reinterpret_cast<float4 *>(row + f) = make_float4(s[f], s[f + 1], s[f + 2], s[f + 3]);
#pragma unroll
            for (int f = P & ~3; f < P; ++f) row[f] = s[f];
            q.info[slot] = make_uint4(__float_as_uint(result), __float_as_uint(aux), meta, p);
        }
        count += c;
        __syncwarp();
    }
}

// Plain coalesced loads (one 128 B line per warp per frame plane): frames that are not 16-byte
// aligned, and the ragged tail of the TMA kernel.
template <int P, int MN, int MD>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, AB_LDG_BLOCKS)
stack_warp_kernel(const __grid_constant__ FrameTable ft, size_t npix, float sigma_low,
                  float sigma_high, unsigned max_iter, CeConst k, float *__restrict__ out,
                  unsigned long long *__restrict__ rejected_total) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpQueue<P> *queues = reinterpret_cast<WarpQueue<P> *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpQueue<P> &q = queues[warp];
    unsigned count = 0;  // parked pixels (warp-uniform)
    unsigned local_rej = 0;

    // npix < 2^32 per launch (checked by the launcher): 32-bit pixel indices
    const unsigned n_groups = (unsigned)((npix + 31) / 32);
    const unsigned warp_global = blockIdx.x * WARPS_PER_BLOCK + warp;
    const unsigned warp_stride = gridDim.x * WARPS_PER_BLOCK;
    for (unsigned g = warp_global; g < n_groups; g += warp_stride) {
        if (count >= (unsigned)DRAIN_AT) {  // the top min(count, 32) slots, one per lane
            const unsigned m = count < 32u ? count : 32u;
            drain_impl<P, true>(q, count - m, m, lane, sigma_low, sigma_high, max_iter, out, local_rej);
            count -= m;
        }
        const unsigned p = g * 32 + lane;
        const bool active = p < npix;
        float s[P];
        float pr[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // NaN iff a sample of the pixel is non-finite
#pragma unroll
        for (int f = 0; f < P; ++f) {
            s[f] = active ? __ldcs(ft.p[f] + p) : 0.0f;
            pr[f & 3] = fmaf(s[f], 0.0f, pr[f & 3]);
        }
        const float probe = (pr[0] + pr[1]) + (pr[2] + pr[3]);
        const int n = drop_non_finite<P>(s, probe);
        process_group<P, MN, MD>(s, n, active, p, lane, q, count, local_rej, sigma_low, sigma_high,
                                 max_iter, k, out);
    }
    if (count) drain_dynamic<P>(q, count, lane, sigma_low, sigma_high, max_iter, out, local_rej);
    const unsigned w = __reduce_add_sync(0xffffffffu, local_rej);
    if (lane == 0 && w) atomicAdd(rejected_total, (unsigned long long)w);
}

// ---------------------------------------------------------------------------------------------
// TMA-staged variant (the production path for 16-byte aligned frames).
//
// Same control structure, but the 32 x P samples of a warp's next pixel group are fetched by the
// TMA engine into a warp-private 4 KiB shared-memory stage while the warp sorts the current group
// out of registers:
//   * frames at a constant stride (one [P, rows, cols] cube): ONE cp.async.bulk.tensor.2d per
//     group (box = 32 pixels x P frames, UTMALDG); otherwise P 128-byte cp.async.bulk copies;
//   * completion on a warp-private mbarrier (complete_tx::bytes); the stage is re-armed by lane 0
//     right after the warp has copied it to registers, so one stage per warp is enough: the fetch
//     of group i+1 overlaps the whole sort of group i;
//   * the loads become conflict-free LDS with immediate offsets: the ~1.5 ALU-pipe instructions
//     of 64-bit address arithmetic per global load disappear from the pipe that bounds the kernel.
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
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

struct TmapParam {
    alignas(64) unsigned char bytes[128];  // CUtensorMap
};

template <int P>
struct WarpTmaSmem {
    alignas(128) float stage[P][32];
    WarpQueue<P> q;
    alignas(8) unsigned long long full;
    unsigned pad[30];
};
// ONE block per SM: measured on B200 at P = 32 (r1), 20 warps at 94 registers beat 24 warps at 80
// (3 x 8, 0.590 ms -> 0.531 ms): the compiler stops rematerialising addresses in the hot loop, and
// the warps of an SM walk adjacent groups.  Frame counts whose per-warp shared memory does not
// allow 20 warps in 227 KiB run as many as fit.
template <int P>
constexpr int tma_warps() {
    constexpr size_t per_warp = sizeof(WarpTmaSmem<P>);
    constexpr size_t fit = (227 * 1024) / per_warp;
    return fit >= AB_TMA_WARPS ? AB_TMA_WARPS : (fit < 1 ? 1 : (int)fit);
}

template <int P, int MN, int MD, bool kTensorMap>
__global__ void __launch_bounds__(tma_warps<P>() * 32, 1)
stack_wtma_kernel(const __grid_constant__ FrameTable ft, const __grid_constant__ TmapParam tmap,
                  unsigned n_groups, float sigma_low, float sigma_high, unsigned max_iter, CeConst k,
                  ab_ctx::StackSlot *__restrict__ slot, ab_ctx::StackSlot *__restrict__ stale_slot,
                  float *__restrict__ out, unsigned long long *__restrict__ rejected_total) {
    unsigned *work_counter = &slot->work;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr unsigned TW = tma_warps<P>();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpTmaSmem<P> &sm = reinterpret_cast<WarpTmaSmem<P> *>(smem_raw)[warp];
    const uint32_t bar = smem_u32(&sm.full);
    const uint32_t stage_addr = smem_u32(&sm.stage[0][0]);
    constexpr uint32_t GROUP_BYTES = P * 32 * sizeof(float);

    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    auto issue = [&](unsigned g) {  // lane 0 only
        mbar_arrive_expect_tx(bar, GROUP_BYTES);
        if constexpr (kTensorMap) {
            tma_load_2d(stage_addr, &tmap, (int)(g * 32), 0, bar);
        } else {
#pragma unroll 4
            for (int f = 0; f < P; ++f)
                tma_load_1d(stage_addr + f * 128, ft.p[f] + (size_t)g * 32, 128, bar);
        }
    };
    // SMs do not run at exactly the same speed: a fully static split left ~6 % of the SM-cycles
    // idle at the tail.  Every warp takes its first `static_rounds` groups strided (no traffic),
    // the last ~1/8 of the groups are handed out through a global counter (one same-address
    // atomic per group for ALL groups measured 2x slower: ~900 atomics/us on one L2 line).  Lane 0
    // fetches the next index one iteration ahead, so the atomic's latency hides behind the sort
    // and the next group's TMA is issued on time.
    const unsigned n_warps = gridDim.x * TW;
    const unsigned static_rounds = (n_groups / n_warps) * AB_STATIC_NUM / AB_STATIC_DEN;
    const unsigned dyn_base = static_rounds * n_warps;
    unsigned g = blockIdx.x * TW + warp;
    unsigned round = 1;  // rounds taken so far (lane 0's view of g_next)
    auto fetch_next = [&](unsigned cur) -> unsigned {  // lane 0 only
        if (round < static_rounds) {
            ++round;
            return cur + n_warps;
        }
        return dyn_base + atomicAdd(work_counter, 1u);
    };
    unsigned g_next = 0;
    if (static_rounds == 0) g = 0xffffffffu;  // tiny launch: everything through the counter
    if (lane == 0) {
        if (static_rounds == 0) g = dyn_base + atomicAdd(work_counter, 1u);
        if (g < n_groups) issue(g);
        g_next = fetch_next(g);
    }
    g = __shfl_sync(0xffffffffu, g, 0);

    unsigned count = 0;  // parked pixels (warp-uniform)
    unsigned local_rej = 0;
    uint32_t parity = 0;
    for (; g < n_groups; parity ^= 1u) {
        if (count >= (unsigned)DRAIN_AT) {  // the top min(count, 32) slots, one per lane
            const unsigned m = count < 32u ? count : 32u;
            drain_impl<P, true>(sm.q, count - m, m, lane, sigma_low, sigma_high, max_iter, out,
                                local_rej);
            count -= m;
        }
        const unsigned p = g * 32 + lane;
        mbar_wait(bar, parity);
        float s[P];
#if AB_PROBE2
        // probe with packed f32x2 FMAs: one FFMA2 per sample pair
        unsigned long long pr2[2] = {0ull, 0ull};
#pragma unroll
        for (int f = 0; f < P; ++f) s[f] = sm.stage[f][lane];
#pragma unroll
        for (int f = 0; f + 1 < P; f += 2) {
            unsigned long long pair;
            asm("mov.b64 %0, {%1, %2};" : "=l"(pair) : "f"(s[f]), "f"(s[f + 1]));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(pr2[(f >> 1) & 1]) : "l"(pair), "l"(0ull));
        }
        float q0, q1, q2, q3;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(pr2[0]));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(q2), "=f"(q3) : "l"(pr2[1]));
        float probe = (q0 + q1) + (q2 + q3);
        if (P & 1) probe = fmaf(s[P - 1], 0.0f, probe);
#else
        float pr[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // NaN iff a sample of the pixel is non-finite
#pragma unroll
        for (int f = 0; f < P; ++f) {
            s[f] = sm.stage[f][lane];
            pr[f & 3] = fmaf(s[f], 0.0f, pr[f & 3]);
        }
        const float probe = (pr[0] + pr[1]) + (pr[2] + pr[3]);
#endif
        // every lane has consumed its column (probe depends on all P loads): refill the stage
        asm volatile("" ::"f"(probe) : "memory");
        __syncwarp();
        // (write-after-read: the loads above have returned their data -- probe depends on all of
        // them -- so the refill needs no proxy fence, only the warp-level ordering)
        const unsigned g_cur_next = __shfl_sync(0xffffffffu, g_next, 0);
        if (lane == 0) {
            if (g_cur_next < n_groups) {
#if AB_PROXY_FENCE
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
                issue(g_cur_next);
                g_next = fetch_next(g_cur_next);  // consumed one iteration later
            } else {
                g_next = 0xffffffffu;
            }
        }
        const int n = drop_non_finite<P>(s, probe);
        process_group<P, MN, MD>(s, n, true, p, lane, sm.q, count, local_rej, sigma_low, sigma_high,
                                 max_iter, k, out);
        g = g_cur_next;
    }
    if (count) drain_dynamic<P>(sm.q, count, lane, sigma_low, sigma_high, max_iter, out, local_rej);
    const unsigned w = __reduce_add_sync(0xffffffffu, local_rej);
    if (lane == 0 && w) atomicAdd(&slot->rejected, (unsigned long long)w);
    // StackResult.rejected_pixels without a memset before the launch: the counts meet in the slot's
    // accumulator (zero since it was cleared 128 launches ago); the block that finishes last
    // publishes the total and block 0 clears the slot the launch after next-but-127 will use
    __syncthreads();
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            stale_slot->work = 0;
            stale_slot->done = 0;
            stale_slot->rejected = 0;
        }
        __threadfence();
        if (atomicAdd(&slot->done, 1u) == gridDim.x - 1) {
            __threadfence();
            *rejected_total = atomicAdd(&slot->rejected, 0ull);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Launchers.  MN/MD = fraction of compare-exchanges in the IMAD form.
template <int P, int MN, int MD>
ab_status launch_ldg(ab_ctx *ctx, const FrameTable &ft, size_t npix, float sl, float sh,
                     unsigned max_iter, float *d_out, unsigned long long *d_rej,
                     cudaStream_t stream) {
    if (npix >= ((size_t)1 << 32))
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "npix %zu >= 2^32 per launch", npix);
    const size_t smem = sizeof(WarpQueue<P>) * WARPS_PER_BLOCK;
    constexpr int threads = WARPS_PER_BLOCK * 32;
    auto kern = stack_warp_kernel<P, MN, MD>;
    AB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    AB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) per_sm = 1;
    const size_t n_blocks_needed = (npix + threads - 1) / threads;
    size_t grid = (size_t)ctx->sm_count * per_sm;  // persistent: one wave, groups strided
    if (grid > n_blocks_needed) grid = n_blocks_needed;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, threads, smem, stream>>>(ft, npix, sl, sh, max_iter, CeConst{1, -1},
                                                    d_out, d_rej);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

// Host side of the tensor map (stack.cu): encodes a [P, npix] f32 view with a box_px x P box.
bool ab_make_stack_tmap(ab_ctx *ctx, const float *base, size_t npix, size_t plane_stride_elems,
                        int P, int box_px, TmapParam *out, int n_rows);

template <int P, int MN, int MD>
ab_status launch_mix(ab_ctx *ctx, const FrameTable &ft, size_t npix, float sl, float sh,
                     unsigned max_iter, float *d_out, unsigned long long *d_rej,
                     cudaStream_t stream) {
    // trailing null frames = padding (ab_stack_launch, 33..63 frames on the next kernel up): the
    // cube has n_real rows, the tensor map fills the box's other rows with NaN
    int n_real = P;
    while (n_real > 1 && !ft.p[n_real - 1]) --n_real;
    const bool padded = n_real < P;
    auto deep = [&](const FrameTable &t, size_t n, float *o) {  // padded mode's way out
        DeepPre pre{};
        return ab_stack_launch_deep(ctx, t, n_real, n, sl, sh, max_iter, pre, o, d_rej, stream);
    };
    bool al = ctx->stack_variant != 2;  // AB_STACK_VARIANT=2: plain loads everywhere
    for (int f = 0; f < n_real; ++f) al = al && ((reinterpret_cast<uintptr_t>(ft.p[f]) & 15) == 0);
    const size_t groups = al && npix < ((size_t)1 << 31) ? npix / 32 : 0;
    if (padded && !groups) {
        AB_CUDA(ctx, cudaMemsetAsync(d_rej, 0, sizeof(unsigned long long), stream));
        return deep(ft, npix, d_out);
    }
    if (groups) {
        bool cube = P > 1 && (!ctx->force_tma1d || padded);
        const ptrdiff_t stride = P > 1 ? ft.p[1] - ft.p[0] : 0;
        for (int f = 1; cube && f < n_real; ++f) cube = (ft.p[f] - ft.p[f - 1]) == stride;
        cube = cube && stride > 0 && (stride % 4) == 0 && (size_t)stride >= npix;
        TmapParam tm{};
        if (cube) cube = ab_make_stack_tmap(ctx, ft.p[0], npix, (size_t)stride, P, 32, &tm, n_real);
        if (padded && !cube) {
            AB_CUDA(ctx, cudaMemsetAsync(d_rej, 0, sizeof(unsigned long long), stream));
            return deep(ft, npix, d_out);
        }
        // scattered frames: P small bulk copies per group measured slower (1.18 ms) than plain
        // coalesced loads (0.92 ms) -> only the tensor-map form is used unless forced
        if (!cube && !ctx->force_tma1d) {
            AB_CUDA(ctx, cudaMemsetAsync(d_rej, 0, sizeof(unsigned long long), stream));
            return launch_ldg<P, MN, MD>(ctx, ft, npix, sl, sh, max_iter, d_out, d_rej, stream);
        }
        constexpr int TW = tma_warps<P>();
        const size_t smem = sizeof(WarpTmaSmem<P>) * TW;
        constexpr int threads = TW * 32;
        auto kern = cube ? stack_wtma_kernel<P, MN, MD, true> : stack_wtma_kernel<P, MN, MD, false>;
        AB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
        int per_sm = 0;
        AB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
        if (per_sm < 1) per_sm = 1;
        size_t grid = (size_t)ctx->sm_count * per_sm;
        const size_t need = (groups + TW - 1) / TW;
        if (grid > need) grid = need;
        const unsigned seq = ctx->work_seq.fetch_add(1);
        ab_ctx::StackSlot *slot = ctx->d_work_counters + (seq % ab_ctx::WORK_COUNTERS);
        ab_ctx::StackSlot *stale = ctx->d_work_counters + ((seq + ab_ctx::WORK_COUNTERS / 2) % ab_ctx::WORK_COUNTERS);
        kern<<<(unsigned)grid, threads, smem, stream>>>(ft, tm, (unsigned)groups, sl, sh, max_iter,
                                                        CeConst{1, -1}, slot, stale, d_out, d_rej);
        AB_LAUNCH_CHECK(ctx);
    } else {
        AB_CUDA(ctx, cudaMemsetAsync(d_rej, 0, sizeof(unsigned long long), stream));
    }
    const size_t done = groups * 32;
    if (padded && done < npix) {  // ragged tail of a padded launch: the deep kernel adds its rejects
        FrameTable tail;
        for (int f = 0; f < AB_MAX_FRAMES; ++f) tail.p[f] = f < n_real ? ft.p[f] + done : nullptr;
        return deep(tail, npix - done, d_out + done);
    }
    if (done < npix) {  // ragged tail (or everything, for unaligned frames): plain loads
        FrameTable tail;
        for (int f = 0; f < AB_MAX_FRAMES; ++f) tail.p[f] = f < P ? ft.p[f] + done : nullptr;
        return launch_ldg<P, MN, MD>(ctx, tail, npix - done, sl, sh, max_iter, d_out + done, d_rej,
                                     stream);
    }
    return AB_OK;
}

// Default mix per frame count (tuned on B200 at P = 32).
constexpr int DEFAULT_MN = 1, DEFAULT_MD = 2;

// Developer builds (-DAB_DEV_VARIANTS, `python -m astroburst_b200.build --variant NAME
// AB_DEV_VARIANTS ...`) also carry alternative mixes for P = 32 (AB_STACK_MIX=1..7) for A/B
// measurements, instantiated in stack_inst_mix*.cu; the shipped library has the default only.
#ifdef AB_DEV_VARIANTS
#define AB_MIX_LIST(X) X(1, 0, 1) X(2, 1, 3) X(3, 2, 5) X(4, 3, 5) X(5, 2, 3) X(6, 3, 4) X(7, 1, 1)
#define AB_MIX_EXTERN(idx, mn, md)                                                               \
    extern template ab_status launch_mix<32, mn, md>(ab_ctx *, const FrameTable &, size_t, float, \
                                                     float, unsigned, float *,                   \
                                                     unsigned long long *, cudaStream_t);
AB_MIX_LIST(AB_MIX_EXTERN)
#undef AB_MIX_EXTERN
#endif

template <int P>
ab_status launch_tile(ab_ctx *ctx, const FrameTable &ft, size_t npix, float sl, float sh,
                      unsigned max_iter, float *d_out, unsigned long long *d_rej,
                      cudaStream_t stream) {
#ifdef AB_DEV_VARIANTS
    if constexpr (P == 32) {
#define AB_MIX_CASE(idx, mn, md)                                                                 \
    case idx:                                                                                    \
        return launch_mix<32, mn, md>(ctx, ft, npix, sl, sh, max_iter, d_out, d_rej, stream);
        switch (ctx->stack_mix) {
            AB_MIX_LIST(AB_MIX_CASE)
            default:
                break;
        }
#undef AB_MIX_CASE
    }
#endif
    return launch_mix<P, DEFAULT_MN, DEFAULT_MD>(ctx, ft, npix, sl, sh, max_iter, d_out, d_rej,
                                                 stream);
}

}  // namespace abk
