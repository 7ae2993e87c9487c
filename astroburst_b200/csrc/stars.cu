// stars.cu -- tile background, star detection and the star-based affine aligner for sm_100a.
//
// Reference semantics (paths relative to src-tauri/src/):
//   math/sigma_clip.rs:4-34                  sigma_clipped_stats (kappa 3, 2 iterations per tile)
//   core/analysis/star_detection.rs:32-84    estimate_background
//   core/analysis/star_detection.rs:86-258   detect_stars
//   core/alignment/affine.rs:24-53           normalize_for_detection
//   core/alignment/affine.rs:129-212         align_channel_affine (+ :214-660 helpers)
//
// Work split.  Everything that touches every pixel runs on the GPU:
//   * normalisation is fused into the loads of the other kernels (never materialised);
//   * background: one block per tile; the tile's exact medians / MADs come from an MSB-first radix
//     select over the tile (valid pixels are positive floats, so their bit patterns are monotone),
//     the sigma-clip `retain` is a shrinking [lo, hi] interval -- no sort, no per-tile buffers;
//   * threshold + ordered stream compaction of the (few) above-threshold pixels;
//   * triangle matching: all ref x tgt descriptor pairs (up to 1.17 G compares), votes in a
//     120 x 120 shared-memory matrix.
// What is left is per-star scalar work on a sparse pixel list (typically < 1 % of the image): the
// 8-connected flood fill in the reference's discovery order, f64 moments, the stable flux sort
// and the 3-px dedup run on the host exactly as the reference does (same summation order, so star
// parameters are bit-identical), like the reference's own serial loop.  RANSAC (2000 x <= 120
// points) is host scalar code too; its stream depends on rayon::current_num_threads() in the
// reference (affine.rs:410), so the thread count is a parameter.
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <math_constants.h>

#include "ab_internal.h"

namespace {

constexpr double MAD_TO_SIGMA = 1.4826;
constexpr int BG_THREADS = 512;

struct Norm {  // normalize_for_detection fused into loads: v' = clamp((v - lo) * inv, 0, 1) as f32
    int on;
    double lo, inv;
};
__device__ __forceinline__ float norm_px(const Norm &nm, float v) {
    if (!nm.on) return v;
    double x = ((double)v - nm.lo) * nm.inv;
    x = x < 0.0 ? 0.0 : x;
    x = x > 1.0 ? 1.0 : x;  // NaN stays NaN (f64::clamp)
    return (float)x;
}

// ------------------------------------------------------------------------- background tiles
// One block per tile runs sigma_clipped_stats (sigma_clip.rs:4-34) as a sequence of exact
// selections over a pixel source.  `OnChip` holds the whole tile (<= 65,536 px, every tile the
// detector makes) in the SM -- 48 values per thread in shared memory plus 16 thread-local at
// 1,024 threads -- so HBM/L2 is read once per tile, and selects with sampled brackets (one pass
// per statistic, see below) with the MSB-first radix select on the float bit patterns as the
// exact way out; `FromGlobal` re-reads global memory (radix select only) and serves caller-chosen
// tile sizes beyond the on-chip capacity (ab_estimate_background).
struct TileCtx {
    const float *img;
    size_t cols;
    size_t r0, c0, th, tw;
    int tw_shift;  // log2(tw) when the tile width is a power of two (the usual 256), else -1
    Norm nm;
};
__device__ __forceinline__ float tile_load(const TileCtx &t, unsigned i) {
    unsigned ir, ic;  // no integer division on the common power-of-two tile width
    if (t.tw_shift >= 0) {
        ir = i >> t.tw_shift;
        ic = i & ((1u << t.tw_shift) - 1u);
    } else {
        ir = i / (unsigned)t.tw;
        ic = i - ir * (unsigned)t.tw;
    }
    return norm_px(t.nm, __ldg(t.img + (t.r0 + ir) * t.cols + (t.c0 + ic)));
}
struct FromGlobal {
    TileCtx t;
    __device__ __forceinline__ int slots() const { return 1; }
    __device__ __forceinline__ bool presampled(float &, float &) const { return false; }
    template <typename Fn>
    __device__ __forceinline__ void for_each(Fn fn) const {
        const unsigned n_px = (unsigned)(t.th * t.tw);
        for (unsigned i = threadIdx.x; i < n_px; i += blockDim.x) fn(tile_load(t, i));
    }
};
constexpr int OC_THREADS = 1024, OC_REG = 16, OC_SMEM = 48;  // 64 slots per thread = 65,536 px
struct OnChip {
    float rv[OC_REG];
    const float *sm;  // [OC_SMEM][OC_THREADS], slot-major: conflict-free
    int n_slots;      // ceil(n_px / OC_THREADS); empty positions hold NaN (never valid)
    float samp0, samp1;  // this thread's two samples for the bracket selection (+inf = none)
    __device__ __forceinline__ void fill(const TileCtx &t, float *smw) {
        const unsigned n_px = (unsigned)(t.th * t.tw);
        n_slots = (int)((n_px + OC_THREADS - 1) / OC_THREADS);
        sm = smw;
        // the sample is taken while the tile streams in: the first valid value at or after two
        // pseudo-random slots of this thread's column (a spatially spread sample, no extra pass)
        const unsigned ns = (unsigned)max(n_slots, 1);
        const unsigned w0 = (threadIdx.x * 2654435761u >> 16) % ns, w1 = (w0 + (ns >> 1)) % ns;
        float p0 = CUDART_INF_F, p1 = CUDART_INF_F;
        auto note = [&](float v, unsigned j) {
            const bool ok = v > 1e-7f && v < CUDART_INF_F;
            if (ok && j >= w0 && p0 == CUDART_INF_F) p0 = v;
            if (ok && j >= w1 && j != w0 && p1 == CUDART_INF_F) p1 = v;
        };
#pragma unroll
        for (int j = 0; j < OC_REG; ++j) {
            const unsigned i = (unsigned)j * OC_THREADS + threadIdx.x;
            rv[j] = i < n_px ? tile_load(t, i) : CUDART_NAN_F;
            note(rv[j], (unsigned)j);
        }
        int j = OC_REG;
        for (; j + 8 <= n_slots; j += 8) {  // 8 loads in flight per thread
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned i = (unsigned)(j + u) * OC_THREADS + threadIdx.x;
                v[u] = i < n_px ? tile_load(t, i) : CUDART_NAN_F;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                smw[(j + u - OC_REG) * OC_THREADS + threadIdx.x] = v[u];
                note(v[u], (unsigned)(j + u));
            }
        }
        for (; j < n_slots; ++j) {
            const unsigned i = (unsigned)j * OC_THREADS + threadIdx.x;
            const float v = i < n_px ? tile_load(t, i) : CUDART_NAN_F;
            smw[(j - OC_REG) * OC_THREADS + threadIdx.x] = v;
            note(v, (unsigned)j);
        }
        samp0 = p0;
        samp1 = ns > 1 ? p1 : CUDART_INF_F;
    }
    __device__ __forceinline__ bool presampled(float &a, float &b) const {
        a = samp0;
        b = samp1;
        return true;
    }
    __device__ __forceinline__ int slots() const { return n_slots; }
    template <typename Fn>
    __device__ __forceinline__ void for_each(Fn fn) const {
#pragma unroll
        for (int j = 0; j < OC_REG; ++j)
            if (j < n_slots) fn(rv[j]);
        const float *p = sm + threadIdx.x;
        int j = OC_REG;
        for (; j + 4 <= n_slots; j += 4, p += 4 * OC_THREADS) {
            const float a = p[0], b = p[OC_THREADS], c = p[2 * OC_THREADS], d = p[3 * OC_THREADS];
            fn(a);
            fn(b);
            fn(c);
            fn(d);
        }
        for (; j < n_slots; ++j, p += OC_THREADS) fn(p[0]);
    }
};
constexpr int FS_SAMPLES = 2048, FS_CAP = 6144;  // sampled-bracket selection (on-chip tiles)
struct Sel {  // shared scratch of one block
    unsigned hist[256];
    unsigned bin, cum;
    unsigned long long total;
    unsigned maxkey;
    // sampled-bracket selection: one value per thread sampled once per tile, its keys sorted per
    // statistic, the keys between two sample quantiles collected in one pass over the tile
    float *samp_v;     // FS_SAMPLES sampled pixel values (NaN = none), null = fast path off
    unsigned *bracket; // FS_CAP keys; the sample's sorted keys live here until the bracket is chosen
    unsigned n_samp, n_tot, n_lo, n_br;
    unsigned *fallbacks;  // optional global counter of statistics that took the radix select
};

// key of one pixel for the current selection, or false when the pixel is not in the set.
// mode 0: value; mode 1: |v - median| as f32 via f64 (sigma_clip.rs:15)
__device__ __forceinline__ bool px_key(float v, int mode, double median, float L, float H,
                                       unsigned &key) {
    if (!((fabsf(v) < CUDART_INF_F) && v > 1e-7f)) return false;
    if (!(v >= L && v <= H)) return false;
    key = __float_as_uint(mode ? (float)fabs((double)v - median) : v);
    return true;
}

template <typename Src>
__device__ __forceinline__ unsigned long long tile_count(const Src &src, Sel &s, float L, float H) {
    if (threadIdx.x == 0) s.total = 0;
    __syncthreads();
    unsigned c = 0;
    src.for_each([&](float v) {
        unsigned key;
        c += px_key(v, 0, 0.0, L, H, key) ? 1u : 0u;
    });
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s.total, (unsigned long long)c);
    __syncthreads();
    const unsigned long long r = s.total;
    __syncthreads();
    return r;
}

// k-th smallest key (0-based); *less = number of keys strictly smaller than the result
template <typename Src>
__device__ __forceinline__ unsigned tile_select(const Src &src, Sel &s, int mode, double median,
                                                float L, float H, unsigned long long k,
                                                unsigned long long *less) {
    unsigned prefix = 0, mask = 0;
    unsigned long long below = 0;
#pragma unroll 1
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (threadIdx.x < 256) s.hist[threadIdx.x] = 0;
        __syncthreads();
        src.for_each([&](float v) {
            unsigned key;
            if (px_key(v, mode, median, L, H, key) && (key & mask) == prefix)
                atomicAdd(&s.hist[(key >> shift) & 255u], 1u);  // (match_any aggregation measured slower)
        });
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long cum = 0;
            unsigned b = 0;
            for (; b < 255; ++b) {
                if (cum + s.hist[b] > k) break;
                cum += s.hist[b];
            }
            s.bin = b;
            s.cum = (unsigned)cum;
        }
        __syncthreads();
        k -= s.cum;
        below += s.cum;
        prefix |= s.bin << shift;
        mask |= 0xffu << shift;
        __syncthreads();
    }
    *less = below;
    return prefix;
}

// largest key strictly below `bound` (exists by construction when called)
template <typename Src>
__device__ __forceinline__ unsigned tile_max_below(const Src &src, Sel &s, int mode, double median,
                                                   float L, float H, unsigned bound) {
    if (threadIdx.x == 0) s.maxkey = 0;
    __syncthreads();
    unsigned m = 0;
    src.for_each([&](float v) {
        unsigned key;
        if (px_key(v, mode, median, L, H, key) && key < bound) m = max(m, key);
    });
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(&s.maxkey, m);
    __syncthreads();
    const unsigned r = s.maxkey;
    __syncthreads();
    return r;
}

// ---- sampled-bracket selection -----------------------------------------------------------------
// A radix select reads the whole tile 4 times per order statistic (and the detector needs 6
// statistics + 3 counts per tile).  Here a sorted 2,048-value sample gives two keys that bracket
// the median with ~99 % probability; ONE pass over the tile counts the keys below the bracket, the
// retained keys in total, and collects the bracket's keys (a few thousand) in shared memory -- each
// warp into its own segment, so the pass has no atomics -- where the exact order statistics are
// selected.  Whenever the bracket misses (or a segment overflows) the caller falls back to the
// radix select, so the result is always exact.
constexpr int FS_SEG = FS_CAP / (OC_THREADS / 32);  // bracket keys per warp
__device__ __forceinline__ unsigned hist_find_warp(const unsigned *hist, unsigned k, unsigned *cum_out) {
    // bin b with cum(b-1) <= k < cum(b) over 256 bins; warp 0 only (all 32 lanes)
    const int lane = threadIdx.x & 31;
    unsigned loc[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        loc[j] = hist[lane * 8 + j];
        sum += loc[j];
    }
    unsigned inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    unsigned run = inc - sum;  // keys in the bins before this lane's
    const bool mine = run <= k && k < inc;
    unsigned b = 0xffffffffu, cum = 0;
    if (mine) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (b == 0xffffffffu && k < run + loc[j]) {
                b = lane * 8 + j;
                cum = run;
            }
            run += (b == 0xffffffffu) ? loc[j] : 0u;
        }
    }
    const unsigned who = __ballot_sync(0xffffffffu, mine);
    const int src = who ? __ffs(who) - 1 : 31;
    b = __shfl_sync(0xffffffffu, b, src);
    cum = __shfl_sync(0xffffffffu, cum, src);
    if (!who) {  // k beyond the total: last bin (callers never ask for it)
        b = 255;
        cum = __shfl_sync(0xffffffffu, inc, 31) - hist[255];
    }
    *cum_out = cum;
    return b;
}
// k-th smallest (0-based) of the bracket (every warp scans its own segment of `mine` keys);
// *less = keys strictly below it.  The bracket's keys lie in [base, base + range]: the radix select
// runs on key - base and starts at the highest byte `range` reaches (typically 2 passes instead of
// 4, and the first pass already spreads the keys over the bins instead of one hot bin).
__device__ __forceinline__ unsigned bracket_select(const unsigned *seg, unsigned mine, Sel &s, unsigned k,
                                                   unsigned base, unsigned range, unsigned *less) {
    unsigned prefix = 0, mask = 0, below = 0;
    const unsigned lane = threadIdx.x & 31;
    const int shift0 = ((31 - __clz((int)(range | 1u))) >> 3) << 3;
#pragma unroll 1
    for (int shift = shift0; shift >= 0; shift -= 8) {
        if (threadIdx.x < 256) s.hist[threadIdx.x] = 0;
        __syncthreads();
        for (unsigned i = lane; i < mine; i += 32) {
            const unsigned key = seg[i] - base;
            if ((key & mask) == prefix) atomicAdd(&s.hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned cum;
            const unsigned b = hist_find_warp(s.hist, k, &cum);
            if (threadIdx.x == 0) {
                s.bin = b;
                s.cum = cum;
            }
        }
        __syncthreads();
        k -= s.cum;
        below += s.cum;
        prefix |= s.bin << shift;
        mask |= 0xffu << shift;
    }
    *less = below;
    return prefix + base;
}
// largest bracket key strictly below `bound` (exists by construction when called)
__device__ __forceinline__ unsigned bracket_max_below(const unsigned *seg, unsigned mine, Sel &s, unsigned bound) {
    __syncthreads();
    if (threadIdx.x == 0) s.maxkey = 0;
    __syncthreads();
    unsigned m = 0;
    for (unsigned i = threadIdx.x & 31; i < mine; i += 32) {
        const unsigned key = seg[i];
        if (key < bound) m = max(m, key);
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&s.maxkey, m);
    __syncthreads();
    return s.maxkey;
}
// ascending bitonic sort of FS_SAMPLES u32 in shared memory, one comparator per thread and stage
__device__ __forceinline__ void sort_samples(unsigned *sk) {
    for (unsigned size = 2; size <= FS_SAMPLES; size <<= 1)
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            const unsigned t = threadIdx.x, p = 2 * t - (t & (stride - 1)), q = p + stride;
            const unsigned a = sk[p], b = sk[q];
            const bool up = (p & size) == 0;
            if ((a > b) == up) {
                sk[p] = b;
                sk[q] = a;
            }
        }
    __syncthreads();
}
// two values per thread from pseudo-random slots of its column (a spatially spread sample), sorted
// by value once per tile: samp_v[0..n_samp) ascending, the rest +inf
template <typename Src>
__device__ __forceinline__ void tile_sample(const Src &src, Sel &s) {
    if (!s.samp_v) return;
    float p0 = CUDART_INF_F, p1 = CUDART_INF_F;
    if (!src.presampled(p0, p1)) {  // (sources that did not sample while loading: one pass)
        const unsigned ns = (unsigned)max(src.slots(), 1);
        const unsigned w0 = (threadIdx.x * 2654435761u >> 16) % ns, w1 = (w0 + (ns >> 1)) % ns;
        unsigned idx = 0;
        src.for_each([&](float v) {
            const bool ok = v > 1e-7f && v < CUDART_INF_F;
            if (ok && idx >= w0 && p0 == CUDART_INF_F) p0 = v;
            if (ok && idx >= w1 && idx != w0 && p1 == CUDART_INF_F) p1 = v;
            ++idx;
        });
        if (ns < 2) p1 = CUDART_INF_F;
    }
    if (threadIdx.x == 0) s.n_samp = 0;
    __syncthreads();
    unsigned c = (p0 < CUDART_INF_F) + (p1 < CUDART_INF_F);
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s.n_samp, c);
    s.samp_v[threadIdx.x] = p0;
    s.samp_v[threadIdx.x + OC_THREADS] = p1;
    sort_samples(reinterpret_cast<unsigned *>(s.samp_v));  // positive floats order like their bits
}
// Fast (left, right, n): returns false when the caller must use the radix select.
template <int MODE, typename Src>
__device__ __forceinline__ bool tile_pair_fast(const Src &src, Sel &s, double median, float L, float H,
                                               unsigned long long *n_out, float *left, float *right) {
    if (!s.samp_v) return false;
    // valid pixel <=> Lq <= v <= Hq (finite, > 1e-7, inside the clip; NaN fails both compares)
    const float Lq = fmaxf(L, __uint_as_float(__float_as_uint(1e-7f) + 1u)), Hq = fminf(H, 3.402823466e38f);
    auto key_of = [&](float v) -> unsigned {
        return __float_as_uint(MODE ? (float)fabs((double)v - median) : v);
    };
    // ---- the bracket [lo_key, hi_key] around the sample's median
    unsigned lo_key = 0u, hi_key = 0xfffffffeu;
    {
        // the retained part of the sorted sample: indices [a, b)
        const unsigned m0 = s.n_samp;
        unsigned a = 0, b = m0;
        if (m0) {
            unsigned l = 0, r = m0;  // first index with value >= Lq
            while (l < r) {
                const unsigned mdl = (l + r) >> 1;
                if (s.samp_v[mdl] < Lq) l = mdl + 1; else r = mdl;
            }
            a = l;
            r = m0;  // first index with value > Hq
            while (l < r) {
                const unsigned mdl = (l + r) >> 1;
                if (s.samp_v[mdl] <= Hq) l = mdl + 1; else r = mdl;
            }
            b = l;
        }
        const unsigned m = b - a;
        const int d = (int)(1.25f * sqrtf((float)m)) + 6;  // ~2.5 sigma of the sample median's rank
        const int i = (int)(m >> 1);
        if (MODE == 0) {
            if (m >= 64) {  // (a thin sample brackets everything: small sets fit the bracket whole)
                if (i - d >= 0) lo_key = __float_as_uint(s.samp_v[a + i - d]);
                if (i + d < (int)m) hi_key = __float_as_uint(s.samp_v[a + i + d]);
            }
            __syncthreads();
        } else {
            // |v - median| over the value-sorted sample is V-shaped: descending up to the first sample
            // >= median (index c), ascending from there.  The k-th smallest key of the two monotone
            // halves comes from a binary search over how many of the k + 1 smallest sit on the left.
            unsigned c = a;
            {
                unsigned l = a, r = b;
                while (l < r) {
                    const unsigned mdl = (l + r) >> 1;
                    if ((double)s.samp_v[mdl] < median) l = mdl + 1; else r = mdl;
                }
                c = l;
            }
            const int nL = (int)(c - a), nR = (int)(b - c);
            auto Lk = [&](int j) { return key_of(s.samp_v[c - 1 - j]); };  // ascending in j
            auto Rk = [&](int j) { return key_of(s.samp_v[c + j]); };
            auto kth = [&](int k) -> unsigned {
                int lo = max(0, k + 1 - nR), hi = min(k + 1, nL);
                while (lo < hi) {
                    const int t = (lo + hi) >> 1;  // t from the left, k + 1 - t from the right
                    if (Lk(t) < Rk(k - t)) lo = t + 1; else hi = t;
                }
                const unsigned x = lo > 0 ? Lk(lo - 1) : 0u, y = k + 1 - lo > 0 ? Rk(k - lo) : 0u;
                return max(x, y);
            };
            if (m >= 64) {
                if (i - d >= 0) lo_key = kth(i - d);
                if (i + d < (int)m) hi_key = kth(i + d);
            }
            __syncthreads();
        }
    }
    // ---- one pass: retained keys in total, keys below the bracket, the bracket's keys
    if (threadIdx.x == 0) s.n_tot = s.n_lo = s.n_br = s.maxkey = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1u;
    unsigned *seg = s.bracket + (threadIdx.x >> 5) * FS_SEG;
    unsigned c_tot = 0, c_lo = 0, mine = 0;  // mine: keys in this warp's segment (warp-uniform)
    src.for_each([&](float v) {
        const bool ok = v >= Lq && v <= Hq;
        const unsigned key = key_of(v);
        const bool in = ok && key >= lo_key && key <= hi_key;
        c_tot += ok ? 1u : 0u;
        c_lo += (ok && key < lo_key) ? 1u : 0u;
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        const unsigned pos = mine + __popc(bal & lt_mask);
        if (in && pos < FS_SEG) seg[pos] = key;
        mine += __popc(bal);
    });
    c_tot = __reduce_add_sync(0xffffffffu, c_tot);
    c_lo = __reduce_add_sync(0xffffffffu, c_lo);
    if (lane == 0) {
        if (c_tot) atomicAdd(&s.n_tot, c_tot);
        if (c_lo) atomicAdd(&s.n_lo, c_lo);
        if (mine) atomicAdd(&s.n_br, mine);
        if (mine > FS_SEG) s.maxkey = 1;  // (flag: a segment overflowed)
    }
    __syncthreads();
    const unsigned n = s.n_tot, n_lo = s.n_lo, n_br = s.n_br, over = s.maxkey;
    __syncthreads();
    *n_out = n;
    if (n == 0) {
        *left = *right = 0.0f;
        return true;
    }
    const unsigned mid = n >> 1, lowest = (n & 1u) ? mid : mid - 1;  // ranks wanted: lowest .. mid
    if (over || lowest < n_lo || mid >= n_lo + n_br) {
        if (threadIdx.x == 0 && s.fallbacks) atomicAdd(s.fallbacks, 1u);
        return false;
    }
    unsigned less;
    const unsigned rk = bracket_select(seg, mine, s, mid - n_lo, lo_key, hi_key - lo_key, &less);
    unsigned lk = rk;
    if (!(n & 1u) && less == mid - n_lo)  // the lower middle is a smaller key: the largest one below
        lk = bracket_max_below(seg, mine, s, rk);
    *left = __uint_as_float(lk);
    *right = __uint_as_float(rk);
    return true;
}

// (left, right) = the (n/2 - 1)-th and (n/2)-th order statistics (left == right for odd n)
template <typename Src>
__device__ __forceinline__ void tile_median_pair(const Src &src, Sel &s, int mode, double median,
                                                 float L, float H, unsigned long long n, float *left,
                                                 float *right) {
    const unsigned long long mid = n / 2;
    unsigned long long less;
    const unsigned rk = tile_select(src, s, mode, median, L, H, mid, &less);
    unsigned lk = rk;
    if (n % 2 == 0 && less == mid) lk = tile_max_below(src, s, mode, median, L, H, rk);
    *left = __uint_as_float(lk);
    *right = __uint_as_float(rk);
}

// mode 0: exact_median_mut (median.rs:27-44); mode 1: the f32 MAD of sigma_clip.rs:14-17.
// *n_io: in = the retained count when known (else ~0), out = the retained count.
template <typename Src>
__device__ __noinline__ double tile_stat(const Src &src, Sel &s, int mode, double median, float L,
                                         float H, unsigned long long *n_io) {
    float l, r;
    unsigned long long n = *n_io;
    const bool fast = mode ? tile_pair_fast<1>(src, s, median, L, H, &n, &l, &r)
                           : tile_pair_fast<0>(src, s, median, L, H, &n, &l, &r);
    if (!fast) {
        if (n == ~0ull) n = tile_count(src, s, L, H);
        if (n > 0) tile_median_pair(src, s, mode, median, L, H, n, &l, &r);
    }
    *n_io = n;
    if (n == 0) return 0.0;
    if (mode == 0) return (n % 2 == 0) ? ((double)l + (double)r) / 2.0 : (double)r;
    const float m = (n % 2 == 0) ? (l + r) / 2.0f : r;  // median_f32_mut
    return (double)m;
}

// sigma_clipped_stats(&mut vals, 3.0, 2) of one tile + the n >= 8 gate of star_detection.rs:61-66.
// The retained count of every clip round comes out of the median's own pass over the tile.
template <typename Src>
__device__ __forceinline__ void tile_background(const Src &src, Sel &s, double *__restrict__ med_out,
                                                double *__restrict__ sig_out, int *__restrict__ ok_out) {
    float L = -CUDART_INF_F, H = CUDART_INF_F;
    tile_sample(src, s);
    unsigned long long n = ~0ull;
    double median = tile_stat(src, s, 0, 0.0, L, H, &n);  // median of everything valid (+ its count)
    if (n < 8) {
        if (threadIdx.x == 0) ok_out[blockIdx.x] = 0;
        return;
    }
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
        if (n < 3) break;
        unsigned long long n_same = n;
        const double mad = tile_stat(src, s, 1, median, L, H, &n_same);
        const double sig = fmax(mad * MAD_TO_SIGMA, 1e-30);
        const float lo = (float)(median - 3.0 * sig);
        const float hi = (float)(median + 3.0 * sig);
        // retain(v >= lo && v <= hi) on the already retained set
        if (!(lo <= L)) L = lo;  // NaN-safe max
        if (!(hi >= H)) H = hi;
        n = ~0ull;
        median = tile_stat(src, s, 0, 0.0, L, H, &n);  // the next round's (or the final) median
    }
    double sigma = 1.0;  // sigma_clip.rs:24-26
    if (n > 0) {
        unsigned long long n_same = n;
        sigma = fmax(tile_stat(src, s, 1, median, L, H, &n_same) * MAD_TO_SIGMA, 1e-30);
    } else {
        median = 0.0;
    }
    if (threadIdx.x == 0) {
        med_out[blockIdx.x] = median;
        sig_out[blockIdx.x] = sigma;
        ok_out[blockIdx.x] = 1;
    }
}

__device__ __forceinline__ TileCtx tile_of_block(const float *img, size_t rows, size_t cols,
                                                 size_t step, size_t tiles_x, const Norm &nm) {
    const size_t ty = blockIdx.x / tiles_x, tx = blockIdx.x % tiles_x;
    TileCtx t;
    t.img = img;
    t.cols = cols;
    t.r0 = ty * step;
    t.c0 = tx * step;
    t.th = min(step, rows - t.r0);
    t.tw = min(step, cols - t.c0);
    t.tw_shift = (t.tw & (t.tw - 1)) == 0 ? __ffsll((long long)t.tw) - 1 : -1;
    t.nm = nm;
    return t;
}

__global__ void __launch_bounds__(BG_THREADS)
bg_tiles_kernel(const float *__restrict__ img, size_t rows, size_t cols, size_t step,
                size_t tiles_x, Norm nm, double *__restrict__ med_out,
                double *__restrict__ sig_out, int *__restrict__ ok_out) {
    __shared__ Sel s;
    if (threadIdx.x == 0) s.samp_v = nullptr;  // tiles beyond the on-chip capacity: radix select only
    __syncthreads();
    FromGlobal src{tile_of_block(img, rows, cols, step, tiles_x, nm)};
    tile_background(src, s, med_out, sig_out, ok_out);
}

__global__ void __launch_bounds__(OC_THREADS, 1)
bg_tiles_onchip_kernel(const float *__restrict__ img, size_t rows, size_t cols, size_t step,
                       size_t tiles_x, Norm nm, double *__restrict__ med_out,
                       double *__restrict__ sig_out, int *__restrict__ ok_out, unsigned *fallbacks) {
    extern __shared__ float oc_smem[];
    __shared__ Sel s;
    const TileCtx t = tile_of_block(img, rows, cols, step, tiles_x, nm);
    OnChip src;
    src.fill(t, oc_smem);
    if (threadIdx.x == 0) {  // fast-path scratch behind the tile's shared-memory part
        const unsigned n_px = (unsigned)(t.th * t.tw);
        const int slots = (int)((n_px + OC_THREADS - 1) / OC_THREADS);
        float *tail = oc_smem + (size_t)(slots > OC_REG ? slots - OC_REG : 0) * OC_THREADS;
        s.samp_v = tail;
        s.bracket = reinterpret_cast<unsigned *>(tail + FS_SAMPLES);
        s.fallbacks = fallbacks;
    }
    __syncthreads();
    tile_background(src, s, med_out, sig_out, ok_out);
}

// step x step tiles of at most 65,536 px run from the on-chip copy
ab_status launch_bg_tiles(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols, size_t step,
                          size_t tiles_x, size_t n_tiles, const Norm &nm, double *d_med,
                          double *d_sig, int *d_ok, cudaStream_t s, unsigned *d_fb = nullptr) {
    if (step * step <= (size_t)OC_THREADS * (OC_REG + OC_SMEM)) {
        const size_t slots = (std::min(step, rows) * std::min(step, cols) + OC_THREADS - 1) / OC_THREADS;
        const size_t smem = (slots > OC_REG ? (slots - OC_REG) * OC_THREADS * sizeof(float) : 0) +
                            (FS_SAMPLES + FS_CAP) * sizeof(unsigned);
        AB_CUDA(ctx, cudaFuncSetAttribute(bg_tiles_onchip_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bg_tiles_onchip_kernel<<<(unsigned)n_tiles, OC_THREADS, smem, s>>>(d_img, rows, cols, step, tiles_x,
                                                                          nm, d_med, d_sig, d_ok, d_fb);
    } else {
        bg_tiles_kernel<<<(unsigned)n_tiles, BG_THREADS, 0, s>>>(d_img, rows, cols, step, tiles_x, nm,
                                                                d_med, d_sig, d_ok);
    }
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

// ------------------------------------------------------------------------- threshold + compaction
// pass 1: per-row count of pixels with (v as f64) > threshold && finite; pass 2: ordered write.
// One block per row, every thread owns RP_PER contiguous pixels of a 256 * RP_PER super-chunk (eight
// float4 loads in flight), keeps them in registers with a 32-bit foreground mask, and -- in pass 2
// -- writes its foreground pixels at row offset + block-wide exclusive scan of the per-thread
// counts: one scan (two barriers) per 8,192 columns instead of one per 256.
// `(v as f64) > thr` is evaluated as `v > thr_dn` with thr_dn = the largest f32 <= thr (floats are
// discrete: the same predicate, NaN false on both sides, no conversion per pixel).
constexpr int RP_PER = 32, RP_THREADS = 256;
__device__ __forceinline__ unsigned row_fg_mask(const float *__restrict__ row, size_t cols, size_t xs, bool vec,
                                                const Norm &nm, float thr_dn, float (&v)[RP_PER]) {
    if (vec && xs + RP_PER <= cols) {
        const float4 *p = reinterpret_cast<const float4 *>(row + xs);
#pragma unroll
        for (int j = 0; j < RP_PER / 4; ++j) {
            const float4 q = __ldcs(p + j);
            v[4 * j] = q.x;
            v[4 * j + 1] = q.y;
            v[4 * j + 2] = q.z;
            v[4 * j + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < RP_PER; ++j) v[j] = xs + j < cols ? __ldcs(row + xs + j) : CUDART_NAN_F;
    }
    unsigned mask = 0;
#pragma unroll
    for (int j = 0; j < RP_PER; ++j) {
        v[j] = norm_px(nm, v[j]);
        mask |= (v[j] > thr_dn && v[j] < CUDART_INF_F) ? (1u << j) : 0u;
    }
    return mask;
}
__global__ void __launch_bounds__(RP_THREADS)
row_count_kernel(const float *__restrict__ img, size_t rows, size_t cols, Norm nm, float thr_dn,
                 unsigned *__restrict__ row_counts) {
    const size_t r = blockIdx.x;
    const float *row = img + r * cols;
    const bool vec = (cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(img) & 15) == 0);
    unsigned c = 0;
    for (size_t x0 = 0; x0 < cols; x0 += (size_t)RP_THREADS * RP_PER) {
        float v[RP_PER];
        c += __popc(row_fg_mask(row, cols, x0 + (size_t)threadIdx.x * RP_PER, vec, nm, thr_dn, v));
    }
    __shared__ unsigned sh;
    if (threadIdx.x == 0) sh = 0;
    __syncthreads();
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&sh, c);
    __syncthreads();
    if (threadIdx.x == 0) row_counts[r] = sh;
}

__global__ void __launch_bounds__(RP_THREADS)
row_compact_kernel(const float *__restrict__ img, size_t rows, size_t cols, Norm nm, float thr_dn,
                   const unsigned long long *__restrict__ row_offsets, unsigned *__restrict__ out_col,
                   float *__restrict__ out_val, unsigned *__restrict__ out_row) {
    const size_t r = blockIdx.x;
    const unsigned long long base = row_offsets[r];
    if (row_offsets[r + 1] == base) return;  // nothing above the threshold in this row
    const float *row = img + r * cols;
    const bool vec = (cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(img) & 15) == 0);
    __shared__ unsigned warp_tot[RP_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned run = 0;  // foreground pixels of the row before this super-chunk (block-uniform)
    for (size_t x0 = 0; x0 < cols; x0 += (size_t)RP_THREADS * RP_PER) {
        float v[RP_PER];
        const size_t xs = x0 + (size_t)threadIdx.x * RP_PER;
        unsigned mask = row_fg_mask(row, cols, xs, vec, nm, thr_dn, v);
        const unsigned cnt = __popc(mask);
        unsigned inc = cnt;  // inclusive scan over the warp, then over the warps
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        __syncthreads();  // (the previous super-chunk's totals have been read)
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        unsigned before = run, total = 0;
#pragma unroll
        for (int w = 0; w < RP_THREADS / 32; ++w) {
            before += w < warp ? warp_tot[w] : 0u;
            total += warp_tot[w];
        }
        unsigned long long pos = base + before + inc - cnt;
#pragma unroll
        for (int j = 0; j < RP_PER; ++j) {
            if (mask & (1u << j)) {
                out_col[pos] = (unsigned)(xs + j);
                out_val[pos] = v[j];
                out_row[pos] = (unsigned)r;
                ++pos;
            }
        }
        run += total;
    }
}

// exclusive prefix sum of the per-row counts -> row offsets (rows + 1 entries), one block
__global__ void __launch_bounds__(1024)
row_offsets_kernel(const unsigned *__restrict__ row_counts, size_t rows,
                   unsigned long long *__restrict__ row_offsets) {
    __shared__ unsigned long long part[1024];
    const size_t per = (rows + blockDim.x - 1) / blockDim.x;
    const size_t b = (size_t)threadIdx.x * per, e = b + per < rows ? b + per : rows;
    unsigned long long sum = 0;
    for (size_t r = b; r < e; ++r) sum += row_counts[r];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (unsigned t = 0; t < blockDim.x; ++t) {
            const unsigned long long v = part[t];
            part[t] = run;
            run += v;
        }
        row_offsets[rows] = run;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (size_t r = b; r < e; ++r) {
        row_offsets[r] = run;
        run += row_counts[r];
    }
}

// strided sample gather for normalize_for_detection (affine.rs:31-32)
__global__ void sample_kernel(const float *__restrict__ img, size_t len, size_t step,
                              float *__restrict__ out, size_t n_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_out) out[i] = img[i * step];
}

// normalize_for_detection materialised (only for the standalone ABI entry / tests)
__global__ void normalize_kernel(const float *__restrict__ img, size_t len, Norm nm,
                                 float *__restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride)
        out[i] = norm_px(nm, img[i]);
}

// ------------------------------------------------------------------------- triangle matching
constexpr int MAX_STARS = 120;  // affine.rs:8
// votes[ri * MAX_STARS + ti] over all (ref tri, tgt tri) pairs within tolerance (affine.rs:327-348).
// 2-D grid: blockIdx.x = 256 reference triangles (one per thread), blockIdx.y = a chunk of target
// triangles staged in shared memory (every thread reads the same entry: broadcast).  Matches are
// rare (tolerance 0.02 on both ratios), so they go straight to the L2-resident vote matrix.
constexpr int TRI_CHUNK = 1024;
__global__ void __launch_bounds__(256)
tri_match_kernel(const double *__restrict__ r_mid, const double *__restrict__ r_long,
                 const unsigned char *__restrict__ r_sorted, int n_ref, const double *__restrict__ t_mid,
                 const double *__restrict__ t_long, const unsigned char *__restrict__ t_sorted,
                 int n_tgt, double tol, unsigned *__restrict__ votes) {
    __shared__ double s_mid[TRI_CHUNK], s_long[TRI_CHUNK];
    __shared__ unsigned char s_srt[3 * TRI_CHUNK];
    const int b0 = blockIdx.y * TRI_CHUNK, nb = min(TRI_CHUNK, n_tgt - b0);
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        s_mid[i] = t_mid[b0 + i];
        s_long[i] = t_long[b0 + i];
    }
    for (int i = threadIdx.x; i < 3 * nb; i += blockDim.x) s_srt[i] = t_sorted[3 * b0 + i];
    __syncthreads();
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_ref) return;
    const double rm = r_mid[a], rl = r_long[a];
    if (!(fabs(rm) < CUDART_INF)) return;  // a triangle the reference drops
    const int r0 = r_sorted[3 * a], r1 = r_sorted[3 * a + 1], r2 = r_sorted[3 * a + 2];
    for (int b = 0; b < nb; ++b) {
        const double d_mid = fabs(rm - s_mid[b]);
        const double d_long = fabs(rl - s_long[b]);
        if (d_mid > tol || d_long > tol) continue;
        atomicAdd(&votes[r0 * MAX_STARS + s_srt[3 * b]], 1u);
        atomicAdd(&votes[r1 * MAX_STARS + s_srt[3 * b + 1]], 1u);
        atomicAdd(&votes[r2 * MAX_STARS + s_srt[3 * b + 2]], 1u);
    }
}

// ------------------------------------------------------------------------- host-side pieces
inline size_t f2us(double x) {  // Rust `f64 as usize`
    if (!(x > 0.0)) return 0;
    if (x >= 18446744073709551615.0) return SIZE_MAX;
    return (size_t)x;
}

// ------------------------------------------------------------------------- components on the device
// star_detection.rs:105-214 on the sparse above-threshold list, bit-identical to the reference's
// serial raster scan + BFS:
//   1. union-find CCL (8-connectivity) over the sparse pixels: every pixel links to its W / NW /
//      N / NE neighbours (one binary search in the previous row), atomicMin hooks, then a flatten
//      pass that also counts npix per root and takes the smallest INTERIOR pixel of the component
//      -- the pixel at which the reference's scan discovers it (seeds are interior pixels,
//      components may contain border pixels, star_detection.rs:108-140);
//   2. components with 3 <= npix <= 5000 (:142-145; larger blobs are flood-filled and dropped by
//      the reference, so only their size matters) get a slot in discovery order (prefix sum over
//      the seed flags) and a slice of the visiting-order queue;
//   3. one 8-lane group per component replays the reference's BFS from the seed: lane j looks up
//      neighbour j of the queue head in the reference's order N, S, W, E, NW, NE, SW, SE, and the
//      unvisited ones are appended in that order (ballot + rank), so the queue IS the reference's
//      `component` vector; the f64 moments are then accumulated over it sequentially, expression
//      by expression as :147-175 (no FMA contraction: -fmad=false), hence bit-identical sums;
//   4. stable radix sort by flux descending (input in discovery order = the reference's stable
//      sort_by, :215).  The greedy 3-px dedup (:217-248) stays host scalar code on the sorted list.
struct CompRec {
    unsigned seed, off, npix, pad;
};
struct StarRec {
    ab_star st;
};

__device__ __forceinline__ unsigned uf_find(const unsigned *parent, unsigned x) {
    for (;;) {
        const unsigned p = __ldcg(parent + x);
        if (p == x) return x;
        x = p;
    }
}
__device__ __forceinline__ void uf_unite(unsigned *parent, unsigned a, unsigned b) {
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a > b) {
            const unsigned t = a;
            a = b;
            b = t;
        }
        const unsigned old = atomicMin(parent + b, a);  // hook the larger root under the smaller
        if (old == b) return;
        b = old;
    }
}
// first list index of row `r` whose column is >= c
__device__ __forceinline__ unsigned long long row_lower_bound(const unsigned *__restrict__ col,
                                                              const unsigned long long *__restrict__ ro,
                                                              size_t r, unsigned c) {
    unsigned long long lo = __ldg(ro + r), hi = __ldg(ro + r + 1);
    while (lo < hi) {
        const unsigned long long mid = (lo + hi) >> 1;
        if (__ldg(col + mid) < c) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void ccl_init_kernel(unsigned *parent, unsigned *npix, unsigned *seed, unsigned char *vis,
                                unsigned *flag, unsigned *size, unsigned n) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        parent[i] = i;
        npix[i] = 0;
        seed[i] = 0xffffffffu;
        vis[i] = 0;
        flag[i] = 0;
        size[i] = 0;
    }
}
__global__ void ccl_union_kernel(const unsigned *__restrict__ col, const unsigned *__restrict__ rowi,
                                 const unsigned long long *__restrict__ ro, unsigned *parent, unsigned n) {
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const unsigned r = __ldg(rowi + p), c = __ldg(col + p);
    if ((unsigned long long)p > __ldg(ro + r) && __ldg(col + p - 1) + 1 == c) uf_unite(parent, p, p - 1);
    if (r > 0) {
        const unsigned long long e = __ldg(ro + r);
        for (unsigned long long q = row_lower_bound(col, ro, r - 1, c ? c - 1 : 0u);
             q < e && __ldg(col + q) <= c + 1; ++q)
            uf_unite(parent, p, (unsigned)q);
    }
}
__global__ void ccl_flatten_kernel(const unsigned *__restrict__ col, const unsigned *__restrict__ rowi,
                                   unsigned *parent, unsigned *npix, unsigned *seed, unsigned n,
                                   unsigned rows, unsigned cols) {
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const unsigned root = uf_find(parent, p);
    parent[p] = root;
    atomicAdd(npix + root, 1u);
    const unsigned r = __ldg(rowi + p), c = __ldg(col + p);
    if (r >= 1 && r + 1 < rows && c >= 1 && c + 1 < cols) atomicMin(seed + root, p);
}
// flag / size at the seed position of every component the reference measures
__global__ void ccl_mark_kernel(const unsigned *__restrict__ parent, const unsigned *__restrict__ npix,
                                const unsigned *__restrict__ seed, unsigned *flag, unsigned *size,
                                unsigned n) {
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || parent[p] != p) return;
    const unsigned sd = seed[p], np = npix[p];
    if (sd != 0xffffffffu && np >= 3 && np <= 5000) {
        flag[sd] = 1;
        size[sd] = np;
    }
}
__global__ void ccl_emit_kernel(const unsigned *__restrict__ flag, const unsigned *__restrict__ size,
                                const unsigned *__restrict__ slot, const unsigned *__restrict__ off,
                                CompRec *comps, unsigned n) {
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n && flag[p]) comps[slot[p]] = {p, off[p], size[p], 0u};
}

// nbr[8 p + j] = index (in the sparse list) of neighbour j of pixel p in the reference's order
// N, S, W, E, NW, NE, SW, SE, or 0xffffffff when that pixel is not above the threshold.  All
// lookups of the BFS replay done up front, fully parallel: the replay itself is a serial walk per
// component, and a step then costs one table read instead of a binary search per lane.
__global__ void __launch_bounds__(256)
ccl_neighbours_kernel(const unsigned *__restrict__ col, const unsigned *__restrict__ rowi,
                      const unsigned long long *__restrict__ ro, unsigned n, unsigned rows, unsigned cols,
                      unsigned *__restrict__ nbr) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned cur = t >> 3, j = t & 7;
    if (cur >= n) return;
    const int dr = (j == 0 || j == 4 || j == 5) ? -1 : ((j == 1 || j == 6 || j == 7) ? 1 : 0);
    const int dc = (j == 2 || j == 4 || j == 6) ? -1 : ((j == 3 || j == 5 || j == 7) ? 1 : 0);
    const unsigned r = __ldg(rowi + cur), c = __ldg(col + cur);
    const long long nr = (long long)r + dr, nc = (long long)c + dc;
    unsigned q = 0xffffffffu;
    if (nr >= 0 && nc >= 0 && nr < (long long)rows && nc < (long long)cols) {
        if (dr == 0) {
            const long long cand = (long long)cur + dc;
            if (cand >= (long long)__ldg(ro + r) && cand < (long long)__ldg(ro + r + 1) &&
                __ldg(col + cand) == (unsigned)nc)
                q = (unsigned)cand;
        } else {
            const unsigned long long lb = row_lower_bound(col, ro, (size_t)nr, (unsigned)nc);
            if (lb < __ldg(ro + nr + 1) && __ldg(col + lb) == (unsigned)nc) q = (unsigned)lb;
        }
    }
    nbr[t] = q;
}

// One 8-lane group per component: BFS replay + moments.  keys[k] = ~bits(flux) (ascending radix
// sort = flux descending), all ones for components the reference drops (:177-179, :199-201).
__global__ void __launch_bounds__(128)
ccl_measure_kernel(const unsigned *__restrict__ col, const unsigned *__restrict__ rowi,
                   const float *__restrict__ val, const unsigned long long *__restrict__ ro,
                   const CompRec *__restrict__ comps, unsigned n_comp, unsigned rows, unsigned cols,
                   double bg_median, double bg_sigma, const unsigned *__restrict__ nbr, unsigned *queue,
                   unsigned char *vis, StarRec *stars, unsigned long long *keys, unsigned *order,
                   unsigned *n_valid) {
    const unsigned lane = threadIdx.x & 31, j = lane & 7, g = lane >> 3;
    const unsigned gmask = 0xffu << (8 * g);
    const unsigned groups = gridDim.x * (blockDim.x >> 3);
    for (unsigned k = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; k < n_comp; k += groups) {
        const CompRec cr = comps[k];
        unsigned *qu = queue + cr.off;
        if (j == 0) {
            qu[0] = cr.seed;
            vis[cr.seed] = 1;
        }
        __syncwarp(gmask);
        unsigned head = 0, tail = 1;
        while (head < tail) {
            const unsigned cur = __ldcg(qu + head);
            const unsigned q = __ldg(nbr + 8 * (size_t)cur + j);
            const bool fresh = q != 0xffffffffu && __ldcg(vis + q) == 0;
            const unsigned bal = (__ballot_sync(gmask, fresh) >> (8 * g)) & 0xffu;
            if (fresh) {
                qu[tail + __popc(bal & ((1u << j) - 1u))] = q;
                vis[q] = 1;
            }
            tail += __popc(bal);
            ++head;
            __syncwarp(gmask);
        }
        // moments in visiting order (every lane of the group computes the same scalars)
        double sum_flux = 0.0, sum_x = 0.0, sum_y = 0.0, peak_val = 0.0;
        for (unsigned i = 0; i < cr.npix; ++i) {
            const unsigned p = __ldcg(qu + i);
            const double w = fmax((double)__ldg(val + p) - bg_median, 0.0);
            sum_flux += w;
            sum_x += (double)__ldg(col + p) * w;
            sum_y += (double)__ldg(rowi + p) * w;
            peak_val = fmax(peak_val, w);
        }
        bool ok = sum_flux > 0.0;
        ab_star st = {};
        if (ok) {
            const double cx = sum_x / sum_flux, cy = sum_y / sum_flux;
            double sum_r2 = 0.0, sum_xx = 0.0, sum_yy = 0.0, sum_xy = 0.0;
            for (unsigned i = 0; i < cr.npix; ++i) {
                const unsigned p = __ldcg(qu + i);
                const double w = fmax((double)__ldg(val + p) - bg_median, 0.0);
                const double dx = (double)__ldg(col + p) - cx;
                const double dy = (double)__ldg(rowi + p) - cy;
                sum_r2 += (dx * dx + dy * dy) * w;
                sum_xx += dx * dx * w;
                sum_yy += dy * dy * w;
                sum_xy += dx * dy * w;
            }
            const double sigma_star = sqrt(sum_r2 / (2.0 * sum_flux));
            const double fwhm = sigma_star * 2.3548200450309493;
            ok = !(fwhm < 0.5 || fwhm > 30.0);
            const double ixx = sum_xx / sum_flux, iyy = sum_yy / sum_flux, ixy = sum_xy / sum_flux;
            const double trace = ixx + iyy;
            const double det = fmax(ixx * iyy - ixy * ixy, 0.0);
            const double disc = sqrt(fmax((trace * trace / 4.0) - det, 0.0));
            const double lambda1 = trace / 2.0 + disc;
            const double lambda2 = fmax(trace / 2.0 - disc, 0.0);
            double ecc = 0.0;
            if (lambda1 > 1e-15) {
                ecc = sqrt(1.0 - lambda2 / lambda1);
                ecc = ecc < 0.0 ? 0.0 : (ecc > 1.0 ? 1.0 : ecc);
            }
            const double snr = bg_sigma <= DBL_EPSILON ? 0.0 : peak_val / bg_sigma;  // confidence.rs:3-8
            st = {cx, cy, sum_flux, fwhm, ecc, peak_val, (uint64_t)cr.npix, snr};
        }
        if (j == 0) {
            stars[k].st = st;
            keys[k] = ok ? ~(unsigned long long)__double_as_longlong(sum_flux) : ~0ull;
            order[k] = k;
            if (ok) atomicAdd(n_valid, 1u);
        }
    }
}
__global__ void ccl_gather_kernel(const StarRec *__restrict__ stars, const unsigned *__restrict__ order,
                                  ab_star *__restrict__ out, unsigned n) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = stars[order[i]].st;
}

// greedy dedup on a 3-px grid (star_detection.rs:217-248) over the flux-sorted list; the grid is a
// flat hash of cell -> chain of kept stars (same decisions as the reference's
// HashMap<(usize, usize), Vec<usize>>)
void dedup_stars(const std::vector<ab_star> &stars, std::vector<ab_star> &out, size_t max_keep = SIZE_MAX) {
    const double dedup_r2 = 9.0, cell = 3.0;
    size_t cap = 16;
    while (cap < stars.size() * 2) cap <<= 1;
    std::vector<long long> head(cap, -1), next(stars.size(), -1);
    std::vector<unsigned long long> cell_of(stars.size());
    auto key_of = [](size_t gy, size_t gx) { return ((unsigned long long)gy << 32) ^ (unsigned long long)gx; };
    auto slot_of = [&](unsigned long long k) { return (size_t)((k * 0x9E3779B97F4A7C15ull) >> 20) & (cap - 1); };
    out.clear();
    for (size_t i = 0; i < stars.size() && out.size() < max_keep; ++i) {
        const ab_star &s = stars[i];
        const size_t gx = f2us(s.x / cell), gy = f2us(s.y / cell);
        bool too_close = false;
        for (size_t ny = (gy ? gy - 1 : 0); ny <= gy + 1 && !too_close; ++ny)
            for (size_t nx = (gx ? gx - 1 : 0); nx <= gx + 1 && !too_close; ++nx) {
                const unsigned long long k = key_of(ny, nx);
                for (long long j = head[slot_of(k)]; j >= 0; j = next[j]) {
                    if (cell_of[j] != k) continue;  // another cell sharing the slot
                    const double dx = s.x - stars[j].x, dy = s.y - stars[j].y;
                    if (dx * dx + dy * dy < dedup_r2) {
                        too_close = true;
                        break;
                    }
                }
            }
        if (!too_close) {
            const unsigned long long k = key_of(gy, gx);
            const size_t sl = slot_of(k);
            cell_of[i] = k;
            next[i] = head[sl];
            head[sl] = (long long)i;
            out.push_back(s);
        }
    }
}

// device scratch of the detector inside ctx scratch slot 1
// max_keep: the caller only needs the first `max_keep` stars of the result (the aligner: 120) -- the
// greedy dedup walks the flux-sorted list in order, so a prefix of it is enough
ab_status detect_on_device(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols, Norm nm,
                           double sigma_threshold, cudaStream_t s, std::vector<ab_star> &stars,
                           double *bg_median_out, double *bg_sigma_out, size_t max_keep = SIZE_MAX) {
    stars.clear();
    if (rows < 3 || cols < 3) {  // star_detection.rs:89-98
        *bg_median_out = 0.0;
        *bg_sigma_out = 1.0;
        return AB_OK;
    }
    if (cols >= ((size_t)1 << 32) || rows >= ((size_t)1 << 31))
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "image too large for the detector");
    const size_t tile_size = std::min(std::max(std::min(rows, cols) / 8, (size_t)32), (size_t)256);
    const size_t step = std::max(tile_size, (size_t)16);
    const size_t tiles_y = (rows + step - 1) / step, tiles_x = (cols + step - 1) / step;
    const size_t n_tiles = tiles_y * tiles_x;

    // scratch slot 1: [med f64 x tiles][sig f64 x tiles][ok int x tiles][row_counts u32 x rows]
    const size_t off_sig = n_tiles * 8, off_ok = off_sig + n_tiles * 8;
    const size_t off_rc = (off_ok + n_tiles * 4 + 255) / 256 * 256;
    const size_t off_ro = (off_rc + rows * 4 + 255) / 256 * 256;
    const size_t off_fb = (off_ro + (rows + 1) * 8 + 255) / 256 * 256;
    const size_t fixed = off_fb + 256;
    ab_status st = ab_scratch_reserve(ctx, 1, fixed);
    if (st != AB_OK) return st;
    char *base = static_cast<char *>(ctx->d_scratch[1]);
    double *d_med = reinterpret_cast<double *>(base), *d_sig = reinterpret_cast<double *>(base + off_sig);
    unsigned *d_fb = reinterpret_cast<unsigned *>(base + off_fb);  // AB_TRACE: radix-select fallbacks
    int *d_ok = reinterpret_cast<int *>(base + off_ok);
    unsigned *d_rc = reinterpret_cast<unsigned *>(base + off_rc);
    unsigned long long *d_ro = reinterpret_cast<unsigned long long *>(base + off_ro);

    // AB_TRACE=1: per-stage wall clock on stderr (tiles / count / compact + D2H / host fill)
    const bool trace = std::getenv("AB_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(now() - t0).count();
    };
    const auto t_start = now();
    if (trace) AB_CUDA(ctx, cudaMemsetAsync(d_fb, 0, 4, s));
    st = launch_bg_tiles(ctx, d_img, rows, cols, step, tiles_x, n_tiles, nm, d_med, d_sig, d_ok, s,
                         trace ? d_fb : nullptr);
    if (st != AB_OK) return st;
    unsigned h_fb = 0;
    if (trace) AB_CUDA(ctx, cudaMemcpyAsync(&h_fb, d_fb, 4, cudaMemcpyDeviceToHost, s));
    std::vector<double> h_med(n_tiles), h_sig(n_tiles);
    std::vector<int> h_ok(n_tiles);
    AB_CUDA(ctx, cudaMemcpyAsync(h_med.data(), d_med, n_tiles * 8, cudaMemcpyDeviceToHost, s));
    AB_CUDA(ctx, cudaMemcpyAsync(h_sig.data(), d_sig, n_tiles * 8, cudaMemcpyDeviceToHost, s));
    AB_CUDA(ctx, cudaMemcpyAsync(h_ok.data(), d_ok, n_tiles * 4, cudaMemcpyDeviceToHost, s));
    AB_CUDA(ctx, cudaStreamSynchronize(s));
    std::vector<double> medians, sigmas;
    for (size_t t = 0; t < n_tiles; ++t)
        if (h_ok[t]) {
            medians.push_back(h_med[t]);
            sigmas.push_back(h_sig[t]);
        }
    double bg_median = 0.0, bg_sigma = 1.0;  // star_detection.rs:70-72
    if (!medians.empty()) {
        std::sort(medians.begin(), medians.end());
        std::sort(sigmas.begin(), sigmas.end());
        bg_median = medians[medians.size() / 2];
        bg_sigma = std::fmax(sigmas[sigmas.size() / 2], 1e-10);
    }
    *bg_median_out = bg_median;
    *bg_sigma_out = bg_sigma;
    const double threshold = bg_median + sigma_threshold * bg_sigma;
    const double t_bg = ms_since(t_start);
    const auto t1 = now();

    // the largest f32 <= threshold: `(v as f64) > threshold`  <=>  `v > thr_dn`
    float thr_dn = (float)threshold;
    if ((double)thr_dn > threshold) thr_dn = std::nextafterf(thr_dn, -INFINITY);
    row_count_kernel<<<(unsigned)rows, RP_THREADS, 0, s>>>(d_img, rows, cols, nm, thr_dn, d_rc);
    AB_LAUNCH_CHECK(ctx);
    row_offsets_kernel<<<1, 1024, 0, s>>>(d_rc, rows, d_ro);
    AB_LAUNCH_CHECK(ctx);
    unsigned long long nfg64 = 0;
    AB_CUDA(ctx, cudaMemcpyAsync(&nfg64, d_ro + rows, 8, cudaMemcpyDeviceToHost, s));
    AB_CUDA(ctx, cudaStreamSynchronize(s));
    const double t_count = ms_since(t1);
    const auto t2 = now();
    if (nfg64 >= (1ull << 29))  // (8 neighbour slots per pixel are indexed with 32 bits)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "detector: %llu pixels above the threshold", nfg64);
    const unsigned nfg = (unsigned)nfg64;
    std::vector<ab_star> sorted;
    unsigned n_comp = 0;
    if (nfg) {
        // scratch slot 2, all sized by nfg: col | val | row | parent | npix | seed | flag | size |
        // slot | off | queue | vis, then (components <= nfg / 3) comps | stars | keys x2 | order x2
        const size_t a4 = ((size_t)nfg * 4 + 255) / 256 * 256;
        const size_t max_comp = (size_t)nfg / 3 + 1;
        const size_t b_comp = (max_comp * sizeof(CompRec) + 255) / 256 * 256;
        const size_t b_star = (max_comp * sizeof(StarRec) + 255) / 256 * 256;
        const size_t b_k8 = (max_comp * 8 + 255) / 256 * 256, b_k4 = (max_comp * 4 + 255) / 256 * 256;
        size_t tmp_scan = 0, tmp_sort = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, (unsigned *)nullptr, (unsigned *)nullptr, (int)nfg, s);
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (unsigned long long *)nullptr,
                                        (unsigned long long *)nullptr, (unsigned *)nullptr,
                                        (unsigned *)nullptr, (int)max_comp, 0, 64, s);
        const size_t b_tmp = (std::max(tmp_scan, tmp_sort) + 255) / 256 * 256;
        const size_t total = 20 * a4 + b_comp + 2 * b_star + 2 * b_k8 + 2 * b_k4 + b_tmp + 256;  // (+ 8 a4: neighbour table)
        st = ab_scratch_reserve(ctx, 2, total);
        if (st != AB_OK) return st;
        char *q = static_cast<char *>(ctx->d_scratch[2]);
        auto take = [&](size_t bytes) {
            char *r = q;
            q += bytes;
            return r;
        };
        unsigned *d_col = (unsigned *)take(a4);
        float *d_val = (float *)take(a4);
        unsigned *d_row = (unsigned *)take(a4), *d_parent = (unsigned *)take(a4);
        unsigned *d_npix = (unsigned *)take(a4), *d_seed = (unsigned *)take(a4);
        unsigned *d_flag = (unsigned *)take(a4), *d_size = (unsigned *)take(a4);
        unsigned *d_slot = (unsigned *)take(a4), *d_off = (unsigned *)take(a4);
        unsigned *d_queue = (unsigned *)take(a4);
        unsigned char *d_vis = (unsigned char *)take(a4);
        unsigned *d_nbr = (unsigned *)take(8 * a4);
        CompRec *d_comps = (CompRec *)take(b_comp);
        StarRec *d_stars = (StarRec *)take(b_star);
        ab_star *d_sorted = (ab_star *)take(b_star);
        unsigned long long *d_keys = (unsigned long long *)take(b_k8), *d_keys2 = (unsigned long long *)take(b_k8);
        unsigned *d_order = (unsigned *)take(b_k4), *d_order2 = (unsigned *)take(b_k4);
        void *d_tmp = take(b_tmp);
        unsigned *d_nvalid = (unsigned *)take(256);

        row_compact_kernel<<<(unsigned)rows, RP_THREADS, 0, s>>>(d_img, rows, cols, nm, thr_dn, d_ro, d_col,
                                                                d_val, d_row);
        AB_LAUNCH_CHECK(ctx);
        const unsigned nb = (nfg + 255) / 256;
        ccl_init_kernel<<<nb, 256, 0, s>>>(d_parent, d_npix, d_seed, d_vis, d_flag, d_size, nfg);
        AB_LAUNCH_CHECK(ctx);
        ccl_union_kernel<<<nb, 256, 0, s>>>(d_col, d_row, d_ro, d_parent, nfg);
        AB_LAUNCH_CHECK(ctx);
        ccl_flatten_kernel<<<nb, 256, 0, s>>>(d_col, d_row, d_parent, d_npix, d_seed, nfg, (unsigned)rows,
                                             (unsigned)cols);
        AB_LAUNCH_CHECK(ctx);
        ccl_mark_kernel<<<nb, 256, 0, s>>>(d_parent, d_npix, d_seed, d_flag, d_size, nfg);
        AB_LAUNCH_CHECK(ctx);
        size_t tb = b_tmp;
        AB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_flag, d_slot, (int)nfg, s));
        tb = b_tmp;
        AB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_size, d_off, (int)nfg, s));
        ctx->launches.fetch_add(2, std::memory_order_relaxed);
        unsigned last[2] = {0, 0};
        AB_CUDA(ctx, cudaMemcpyAsync(&last[0], d_slot + (nfg - 1), 4, cudaMemcpyDeviceToHost, s));
        AB_CUDA(ctx, cudaMemcpyAsync(&last[1], d_flag + (nfg - 1), 4, cudaMemcpyDeviceToHost, s));
        AB_CUDA(ctx, cudaMemsetAsync(d_nvalid, 0, 4, s));
        AB_CUDA(ctx, cudaStreamSynchronize(s));
        n_comp = last[0] + last[1];
        if (n_comp) {
            ccl_emit_kernel<<<nb, 256, 0, s>>>(d_flag, d_size, d_slot, d_off, d_comps, nfg);
            AB_LAUNCH_CHECK(ctx);
            ccl_neighbours_kernel<<<(unsigned)(((size_t)nfg * 8 + 255) / 256), 256, 0, s>>>(
                d_col, d_row, d_ro, nfg, (unsigned)rows, (unsigned)cols, d_nbr);
            AB_LAUNCH_CHECK(ctx);
            const unsigned mb = (unsigned)std::min<size_t>(((size_t)n_comp + 15) / 16, (size_t)ctx->sm_count * 16);
            ccl_measure_kernel<<<mb, 128, 0, s>>>(d_col, d_row, d_val, d_ro, d_comps, n_comp, (unsigned)rows,
                                                 (unsigned)cols, bg_median, bg_sigma, d_nbr, d_queue, d_vis,
                                                 d_stars, d_keys, d_order, d_nvalid);
            AB_LAUNCH_CHECK(ctx);
            tb = b_tmp;
            AB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_keys, d_keys2, d_order, d_order2,
                                                         (int)n_comp, 0, 64, s));
            ctx->launches.fetch_add(1, std::memory_order_relaxed);
            unsigned n_valid = 0;
            AB_CUDA(ctx, cudaMemcpyAsync(&n_valid, d_nvalid, 4, cudaMemcpyDeviceToHost, s));
            AB_CUDA(ctx, cudaStreamSynchronize(s));
            if (n_valid) {
                ccl_gather_kernel<<<(n_valid + 255) / 256, 256, 0, s>>>(d_stars, d_order2, d_sorted, n_valid);
                AB_LAUNCH_CHECK(ctx);
                // the whole list, or (max_keep given) prefixes of it until the dedup has enough
                size_t have = 0, want = max_keep == SIZE_MAX ? n_valid : std::min<size_t>(n_valid, 8 * max_keep + 64);
                for (;;) {
                    sorted.resize(want);
                    AB_CUDA(ctx, cudaMemcpyAsync(sorted.data() + have, d_sorted + have, (want - have) * sizeof(ab_star),
                                                 cudaMemcpyDeviceToHost, s));
                    AB_CUDA(ctx, cudaStreamSynchronize(s));
                    have = want;
                    if (have == n_valid) break;
                    dedup_stars(sorted, stars, max_keep);
                    if (stars.size() >= max_keep) break;
                    want = std::min<size_t>(n_valid, have * 4);
                }
            }
        }
    }
    const double t_ccl = ms_since(t2);
    const auto t3 = now();
    dedup_stars(sorted, stars, max_keep);
    if (trace)
        fprintf(stderr, "[ab] detect %zux%zu: tiles %.2f ms, count %.2f ms, compact+CCL+measure+sort %.2f ms "
                        "(%u px, %u components), host dedup %.2f ms (%zu stars); %u of %zu tile statistics by radix select\n",
                rows, cols, t_bg, t_count, t_ccl, nfg, n_comp, ms_since(t3), stars.size(), h_fb, n_tiles * 6);
    return AB_OK;
}

// affine.rs:24-53: percentiles of a strided sample -> (lo, inv_range); on=0 when the reference
// returns the image unchanged
// the two ranks of the sorted finite samples (affine.rs:36-41); non-finite samples were replaced by
// +inf before the sort, *n_fin counts the finite ones
__global__ void sample_ranks_kernel(const float *__restrict__ sorted, const unsigned *__restrict__ n_fin,
                                    float *__restrict__ out2) {
    const size_t n = *n_fin;
    if (n < 100) return;
    out2[0] = sorted[n / 100];
    out2[1] = sorted[n * 999 / 1000];
}
__global__ void sample_finite_kernel(const float *__restrict__ img, size_t step, float *__restrict__ out,
                                     size_t n_out, unsigned *__restrict__ n_fin) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool fin = false;
    if (i < n_out) {
        const float v = img[i * step];
        fin = fabsf(v) < CUDART_INF_F;
        out[i] = fin ? v : CUDART_INF_F;
    }
    const unsigned c = __popc(__ballot_sync(0xffffffffu, fin));
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(n_fin, c);
}
ab_status normalize_params(ab_ctx *ctx, const float *d_img, size_t len, cudaStream_t s, Norm *nm) {
    nm->on = 0;
    nm->lo = 0.0;
    nm->inv = 1.0;
    if (len == 0) return AB_OK;
    const size_t step = std::max(len / 100000, (size_t)1);
    const size_t n_s = (len + step - 1) / step;
    if (n_s >= ((size_t)1 << 31)) return ab_fail(ctx, AB_ERR_UNSUPPORTED, "normalize_for_detection: too many samples");
    // strided sample, finite filter, radix sort and the two rank reads all on the device: one
    // 16-byte D2H instead of the whole sample and two host selections
    size_t tmp_sort = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_sort, (float *)nullptr, (float *)nullptr, (int)n_s, 0, 32, s);
    const size_t b_keys = (n_s * 4 + 255) / 256 * 256, b_tmp = (tmp_sort + 255) / 256 * 256;
    ab_status st = ab_scratch_reserve(ctx, 2, 2 * b_keys + b_tmp + 512);
    if (st != AB_OK) return st;
    char *base = static_cast<char *>(ctx->d_scratch[2]);
    float *d_s = reinterpret_cast<float *>(base), *d_sorted = reinterpret_cast<float *>(base + b_keys);
    void *d_tmp = base + 2 * b_keys;
    unsigned *d_nfin = reinterpret_cast<unsigned *>(base + 2 * b_keys + b_tmp);
    float *d_out2 = reinterpret_cast<float *>(d_nfin + 4);
    AB_CUDA(ctx, cudaMemsetAsync(d_nfin, 0, 64, s));
    sample_finite_kernel<<<(unsigned)((n_s + 255) / 256), 256, 0, s>>>(d_img, step, d_s, n_s, d_nfin);
    AB_LAUNCH_CHECK(ctx);
    size_t tb = b_tmp;
    AB_CUDA(ctx, cub::DeviceRadixSort::SortKeys(d_tmp, tb, d_s, d_sorted, (int)n_s, 0, 32, s));
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    sample_ranks_kernel<<<1, 1, 0, s>>>(d_sorted, d_nfin, d_out2);
    AB_LAUNCH_CHECK(ctx);
    struct {
        unsigned n_fin, pad[3];
        float v[2];
    } h;
    AB_CUDA(ctx, cudaMemcpyAsync(&h, d_nfin, sizeof(h), cudaMemcpyDeviceToHost, s));
    AB_CUDA(ctx, cudaStreamSynchronize(s));
    if (h.n_fin < 100) return AB_OK;
    const double lo = (double)h.v[0];
    const double hi = (double)h.v[1];
    const double range = hi - lo;
    if (range < 1e-15) return AB_OK;
    nm->on = 1;
    nm->lo = lo;
    nm->inv = 1.0 / range;
    return AB_OK;
}

// ---- aligner host math (affine.rs:272-660), f64 scalar code
struct Pt {
    double x, y;
};
struct Match {
    double rx, ry, tx, ty;
};
constexpr double TRIANGLE_TOLERANCE = 0.02;
constexpr size_t MIN_MATCHES_AFFINE = 6, MIN_MATCHES_RIGID = 4, RANSAC_ITERATIONS = 2000;
constexpr double RANSAC_INLIER_PX = 3.0, DETECTION_SIGMA = 3.5, MIN_TRIANGLE_SIDE = 15.0;
constexpr double MIN_INLIER_RATIO = 0.20, MAX_RESIDUAL_PX = 5.0, MAX_OFFSET_FRACTION = 0.40;
constexpr double MAX_ROTATION_DEG = 30.0, MIN_SCALE = 0.70, MAX_SCALE = 1.40;

ab_affine aff_identity() { return {1.0, 0.0, 0.0, 0.0, 1.0, 0.0}; }
__host__ __device__ inline void amap(const ab_affine &t, double x, double y, double *px, double *py) {
    *px = t.a * x + t.b * y + t.tx;
    *py = t.c * x + t.d * y + t.ty;
}
__host__ __device__ inline bool solve_3x3(const double a[3][3], const double b[3], double x[3]) {
    const double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) -
                       a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                       a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    if (fabs(det) < 1e-12) return false;
    const double id = 1.0 / det;
    const double inv[3][3] = {{(a[1][1] * a[2][2] - a[1][2] * a[2][1]) * id,
                               (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id,
                               (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id},
                              {(a[1][2] * a[2][0] - a[1][0] * a[2][2]) * id,
                               (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id,
                               (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id},
                              {(a[1][0] * a[2][1] - a[1][1] * a[2][0]) * id,
                               (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id,
                               (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id}};
    for (int i = 0; i < 3; ++i) x[i] = inv[i][0] * b[0] + inv[i][1] * b[1] + inv[i][2] * b[2];
    return true;
}
__host__ __device__ inline bool solve_ls(const Match *m, size_t n, bool solve_x, double x[3]) {
    double ata[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, atb[3] = {0, 0, 0};
    for (size_t q = 0; q < n; ++q) {
        const double target = solve_x ? m[q].tx : m[q].ty;
        const double row[3] = {m[q].rx, m[q].ry, 1.0};
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) ata[i][j] += row[i] * row[j];
            atb[i] += row[i] * target;
        }
    }
    return solve_3x3(ata, atb, x);
}
__host__ __device__ inline bool fit_affine(const Match *m, size_t n, ab_affine *t) {
    if (n < 3) return false;
    double ab[3], cd[3];
    if (!solve_ls(m, n, true, ab) || !solve_ls(m, n, false, cd)) return false;
    *t = {ab[0], ab[1], ab[2], cd[0], cd[1], cd[2]};
    return true;
}
bool fit_rigid(const Match *m, size_t n, ab_affine *t) {
    if (n < 2) return false;
    double rcx = 0, rcy = 0, tcx = 0, tcy = 0;
    for (size_t q = 0; q < n; ++q) {
        rcx += m[q].rx;
        rcy += m[q].ry;
        tcx += m[q].tx;
        tcy += m[q].ty;
    }
    const double nf = (double)n;
    rcx /= nf;
    rcy /= nf;
    tcx /= nf;
    tcy /= nf;
    double num = 0.0, den = 0.0;
    for (size_t q = 0; q < n; ++q) {
        const double drx = m[q].rx - rcx, dry = m[q].ry - rcy, dtx = m[q].tx - tcx, dty = m[q].ty - tcy;
        num += drx * dty - dry * dtx;
        den += drx * dtx + dry * dty;
    }
    const double theta = std::atan2(num, den), ct = std::cos(theta), sn = std::sin(theta);
    *t = {ct, -sn, tcx - ct * rcx + sn * rcy, sn, ct, tcy - sn * rcx - ct * rcy};
    return true;
}
double residual(const std::vector<Match> &m, const ab_affine &t) {
    if (m.empty()) return 0.0;
    double sum = 0.0;
    for (const Match &q : m) {
        double px, py;
        amap(t, q.rx, q.ry, &px, &py);
        const double ex = px - q.tx, ey = py - q.ty;
        sum += std::sqrt(ex * ex + ey * ey);
    }
    return sum / (double)m.size();
}
bool ransac(const std::vector<Match> &matches, int method, int n_threads_in, ab_align_result *res) {
    const size_t n = matches.size();
    const size_t min_sample = method == 0 ? 3 : 2;
    if (n < min_sample) return false;
    const size_t num_threads = (size_t)std::max(n_threads_in, 1);
    const size_t chunk = (RANSAC_ITERATIONS + num_threads - 1) / num_threads;
    size_t best_inl = 0;
    ab_affine best_t = aff_identity();
    std::vector<char> best_mask(n, 0);
    for (size_t tid = 0; tid < num_threads; ++tid) {  // rayon reduce_with keeps the first maximum
        uint64_t state = 0xDEADBEEFCAFEBABEull + (uint64_t)tid * 0x9E3779B97F4A7C15ull;
        auto rnd = [&]() -> size_t {
            state ^= state << 13;
            state ^= state >> 7;
            state ^= state << 17;
            return (size_t)(state % (uint64_t)n);
        };
        size_t lbest = 0;
        ab_affine lt = aff_identity();
        std::vector<char> lmask(n, 0), mask(n, 0);
        for (size_t it = 0; it < chunk; ++it) {
            size_t sample[3], ns = 0;
            int attempts = 0;
            while (ns < min_sample && attempts < 20) {
                const size_t idx = rnd();
                bool has = false;
                for (size_t q = 0; q < ns; ++q) has |= sample[q] == idx;
                if (!has) sample[ns++] = idx;
                ++attempts;
            }
            if (ns < min_sample) continue;
            Match sm[3];
            for (size_t q = 0; q < ns; ++q) sm[q] = matches[sample[q]];
            ab_affine t;
            if (!(method == 0 ? fit_affine(sm, ns, &t) : fit_rigid(sm, ns, &t))) continue;
            std::fill(mask.begin(), mask.end(), 0);
            size_t inl = 0;
            for (size_t i = 0; i < n; ++i) {
                double px, py;
                amap(t, matches[i].rx, matches[i].ry, &px, &py);
                const double ex = px - matches[i].tx, ey = py - matches[i].ty;
                if (std::sqrt(ex * ex + ey * ey) < RANSAC_INLIER_PX) {
                    ++inl;
                    mask[i] = 1;
                }
            }
            if (inl > lbest) {
                lbest = inl;
                lt = t;
                lmask = mask;
            }
        }
        if (tid == 0 || lbest > best_inl) {
            best_inl = lbest;
            best_t = lt;
            best_mask = lmask;
        }
    }
    if (best_inl < MIN_MATCHES_RIGID) return false;
    if ((double)best_inl / (double)n < MIN_INLIER_RATIO) return false;
    std::vector<Match> inl;
    for (size_t i = 0; i < n; ++i)
        if (best_mask[i]) inl.push_back(matches[i]);
    ab_affine refined;
    if (!(method == 0 ? fit_affine(inl.data(), inl.size(), &refined)
                      : fit_rigid(inl.data(), inl.size(), &refined)))
        refined = best_t;
    const double res_px = residual(inl, refined);
    if (res_px > MAX_RESIDUAL_PX) return false;
    res->transform = refined;
    res->matched_stars = n;
    res->inliers = best_inl;
    res->residual_px = res_px;
    res->method = method;
    return true;
}
// ---- ransac_affine, method "affine", on the device ---------------------------------------------
// The hypotheses are independent: one thread per RANSAC iteration fits its 3-match sample and
// counts inliers; atomicMax on (inliers, first iteration wins ties) is exactly the reference's
// "strictly greater replaces" scan over the chunks in order (affine.rs:426-478); one thread then
// refits on the winner's inliers and takes the residual.  Only + - * / sqrt in f64 (no FMA
// contraction: -fmad=false), so every number equals the host code's.  The xorshift sample indices
// come from the host (a few thousand integer draws, sequential per chunk by construction).
// The rigid model (atan2 / sin / cos, whose device versions round differently) stays host code.
struct RansacOut {
    ab_affine refined;
    double residual;
    unsigned long long inliers;
};
__device__ __forceinline__ bool is_inlier(const ab_affine &t, const Match &q) {
    double px, py;
    amap(t, q.rx, q.ry, &px, &py);
    const double ex = px - q.tx, ey = py - q.ty;
    return sqrt(ex * ex + ey * ey) < RANSAC_INLIER_PX;
}
__global__ void __launch_bounds__(128)
ransac_hypotheses_kernel(const Match *__restrict__ m, int n, const unsigned short *__restrict__ samples,
                         int n_iter, unsigned long long *__restrict__ best_key) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_iter) return;
    const unsigned short s0 = samples[3 * i], s1 = samples[3 * i + 1], s2 = samples[3 * i + 2];
    if (s0 == 0xffffu) return;  // no 3 distinct draws in 20 attempts: the iteration is skipped
    const Match sm[3] = {m[s0], m[s1], m[s2]};
    ab_affine t;
    if (!fit_affine(sm, 3, &t)) return;
    unsigned inl = 0;
    for (int q = 0; q < n; ++q) inl += is_inlier(t, m[q]) ? 1u : 0u;
    if (inl) atomicMax(best_key, ((unsigned long long)inl << 32) | (0xffffffffu - (unsigned)i));
}
__global__ void ransac_refine_kernel(const Match *__restrict__ m, int n, const unsigned short *__restrict__ samples,
                                     const unsigned long long *__restrict__ best_key, RansacOut *__restrict__ out) {
    const unsigned long long key = *best_key;
    out->inliers = key >> 32;
    out->residual = 0.0;
    out->refined = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0};
    if (key == 0) return;
    const unsigned i = 0xffffffffu - (unsigned)(key & 0xffffffffu);
    const Match sm[3] = {m[samples[3 * i]], m[samples[3 * i + 1]], m[samples[3 * i + 2]]};
    ab_affine best;
    fit_affine(sm, 3, &best);
    // fit_affine over the inliers in list order (solve_ls accumulates row by row)
    ab_affine refined = best;
    double xs[2][3];
    bool ok = (key >> 32) >= 3;
    for (int pass = 0; pass < 2 && ok; ++pass) {
        double ata[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, atb[3] = {0, 0, 0};
        for (int q = 0; q < n; ++q) {
            if (!is_inlier(best, m[q])) continue;
            const double target = pass == 0 ? m[q].tx : m[q].ty;
            const double row[3] = {m[q].rx, m[q].ry, 1.0};
            for (int a = 0; a < 3; ++a) {
                for (int b = 0; b < 3; ++b) ata[a][b] += row[a] * row[b];
                atb[a] += row[a] * target;
            }
        }
        ok = solve_3x3(ata, atb, xs[pass]);
    }
    if (ok) refined = {xs[0][0], xs[0][1], xs[0][2], xs[1][0], xs[1][1], xs[1][2]};
    double sum = 0.0;
    unsigned cnt = 0;
    for (int q = 0; q < n; ++q) {
        if (!is_inlier(best, m[q])) continue;
        double px, py;
        amap(refined, m[q].rx, m[q].ry, &px, &py);
        const double ex = px - m[q].tx, ey = py - m[q].ty;
        sum += sqrt(ex * ex + ey * ey);
        ++cnt;
    }
    out->refined = refined;
    out->residual = cnt ? sum / (double)cnt : 0.0;
}
// method 0 of ransac() with the hypotheses on the device; *found as ransac()'s return value
ab_status ransac_affine_device(ab_ctx *ctx, const std::vector<Match> &matches, int n_threads_in, cudaStream_t s,
                               ab_align_result *res, bool *found) {
    *found = false;
    const size_t n = matches.size();
    if (n < 3) return AB_OK;
    if (n >= 0xffffu) {  // (sample indices travel as u16; star lists hold <= 120 matches)
        *found = ransac(matches, 0, n_threads_in, res);
        return AB_OK;
    }
    const size_t num_threads = (size_t)std::max(n_threads_in, 1);
    const size_t chunk = (RANSAC_ITERATIONS + num_threads - 1) / num_threads;
    const size_t n_iter = chunk * num_threads;
    std::vector<unsigned short> samples(3 * n_iter);
    for (size_t tid = 0; tid < num_threads; ++tid) {  // the draws of ransac(), nothing else
        uint64_t state = 0xDEADBEEFCAFEBABEull + (uint64_t)tid * 0x9E3779B97F4A7C15ull;
        for (size_t it = 0; it < chunk; ++it) {
            size_t sample[3], ns = 0;
            int attempts = 0;
            while (ns < 3 && attempts < 20) {
                state ^= state << 13;
                state ^= state >> 7;
                state ^= state << 17;
                const size_t idx = (size_t)(state % (uint64_t)n);
                bool has = false;
                for (size_t q = 0; q < ns; ++q) has |= sample[q] == idx;
                if (!has) sample[ns++] = idx;
                ++attempts;
            }
            unsigned short *o = &samples[3 * (tid * chunk + it)];
            for (int q = 0; q < 3; ++q) o[q] = ns == 3 ? (unsigned short)sample[q] : (unsigned short)0xffffu;
        }
    }
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t o_m = 0, o_s = al(n * sizeof(Match)), o_k = o_s + al(samples.size() * 2), o_o = o_k + 256;
    ab_status st = ab_scratch_reserve(ctx, 2, o_o + sizeof(RansacOut));  // (the matcher's slot: its data is dead)
    if (st != AB_OK) return st;
    char *b = static_cast<char *>(ctx->d_scratch[2]);
    AB_CUDA(ctx, cudaMemcpyAsync(b + o_m, matches.data(), n * sizeof(Match), cudaMemcpyHostToDevice, s));
    AB_CUDA(ctx, cudaMemcpyAsync(b + o_s, samples.data(), samples.size() * 2, cudaMemcpyHostToDevice, s));
    AB_CUDA(ctx, cudaMemsetAsync(b + o_k, 0, 8, s));
    const Match *d_m = reinterpret_cast<const Match *>(b + o_m);
    const unsigned short *d_s = reinterpret_cast<const unsigned short *>(b + o_s);
    unsigned long long *d_k = reinterpret_cast<unsigned long long *>(b + o_k);
    ransac_hypotheses_kernel<<<(unsigned)((n_iter + 127) / 128), 128, 0, s>>>(d_m, (int)n, d_s, (int)n_iter, d_k);
    AB_LAUNCH_CHECK(ctx);
    ransac_refine_kernel<<<1, 1, 0, s>>>(d_m, (int)n, d_s, d_k, reinterpret_cast<RansacOut *>(b + o_o));
    AB_LAUNCH_CHECK(ctx);
    RansacOut h;
    AB_CUDA(ctx, cudaMemcpyAsync(&h, b + o_o, sizeof(h), cudaMemcpyDeviceToHost, s));
    AB_CUDA(ctx, cudaStreamSynchronize(s));
    if (h.inliers < MIN_MATCHES_RIGID) return AB_OK;
    if ((double)h.inliers / (double)n < MIN_INLIER_RATIO) return AB_OK;
    if (h.residual > MAX_RESIDUAL_PX) return AB_OK;
    res->transform = h.refined;
    res->matched_stars = n;
    res->inliers = h.inliers;
    res->residual_px = h.residual;
    res->method = 0;
    *found = true;
    return AB_OK;
}
bool sane(const ab_align_result &r, size_t rows, size_t cols) {
    const ab_affine &t = r.transform;
    if (std::fabs(t.tx) > (double)cols * MAX_OFFSET_FRACTION ||
        std::fabs(t.ty) > (double)rows * MAX_OFFSET_FRACTION)
        return false;
    if (std::fabs(std::atan2(t.c, t.a) * (180.0 / M_PI)) > MAX_ROTATION_DEG) return false;
    const double sx = std::sqrt(t.a * t.a + t.c * t.c), sy = std::sqrt(t.b * t.b + t.d * t.d);
    return !(sx < MIN_SCALE || sx > MAX_SCALE || sy < MIN_SCALE || sy > MAX_SCALE);
}
// fallback_phase_correlation, affine.rs:243-270 (on the original, un-normalised planes)
ab_status fallback(ab_ctx *ctx, const float *d_ref, const float *d_tgt, size_t rows, size_t cols,
                   cudaStream_t s, ab_align_result *out) {
    out->transform = aff_identity();
    out->matched_stars = 0;
    out->inliers = 0;
    out->residual_px = 0.0;
    out->method = 3;
    if (!d_ref) return AB_OK;  // no reference plane on this GPU
    ab_phase_result pc;
    ab_status st = ab_phase_correlate_on(ctx, d_ref, d_tgt, rows, cols, s, &pc);
    if (st != AB_OK) return st;
    const double max_tx = (double)cols * MAX_OFFSET_FRACTION, max_ty = (double)rows * MAX_OFFSET_FRACTION;
    if (std::fabs(pc.dx) > max_tx || std::fabs(pc.dy) > max_ty || pc.confidence < 1.5) return AB_OK;
    out->transform.tx = pc.dx;  // AffineTransform::translation(pc.dx, pc.dy)
    out->transform.ty = pc.dy;
    out->method = 2;
    return AB_OK;
}

// build_triangles + the vertex ordering of match_triangles (affine.rs:279-318, :332-346) for one star
// list: thread t = the t-th triple i < j < k of the first 60 stars in colexicographic order
// (t = C(k,3) + C(j,2) + i).  Triangles the reference drops (shortest side < 15 px, or a vertex
// beyond the list) get descriptor `bad` = +inf (reference list) / -inf (target list): |+-inf - x|
// and |inf - (-inf)| are +inf > tolerance, so they never vote.  Votes are sums, so the order in
// which the triangles sit in the arrays does not matter.
constexpr int TRI_LIMIT = 60, TRI_MAX = TRI_LIMIT * (TRI_LIMIT - 1) * (TRI_LIMIT - 2) / 6;  // 34,220
__device__ __forceinline__ double dev_pdist(double ax, double ay, double bx, double by) {
    const double dx = ax - bx, dy = ay - by;
    return sqrt(dx * dx + dy * dy);
}
__global__ void build_tri_kernel(const double *__restrict__ xy, int n_stars, double bad, double min_side,
                                 double *__restrict__ mid, double *__restrict__ lng,
                                 unsigned char *__restrict__ srt) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= TRI_MAX) return;
    int k = 2;
    while ((k + 1) * k * (k - 1) / 6 <= t) ++k;  // largest k with C(k,3) <= t
    int rem = t - k * (k - 1) * (k - 2) / 6;
    int j = 1;
    while ((j + 1) * j / 2 <= rem) ++j;          // largest j with C(j,2) <= rem
    const int i = rem - j * (j - 1) / 2;
    const int limit = n_stars < TRI_LIMIT ? n_stars : TRI_LIMIT;
    double m = bad, l = bad;
    unsigned char v0 = 0, v1 = 0, v2 = 0;
    if (k < limit) {
        const double xi = xy[2 * i], yi = xy[2 * i + 1], xj = xy[2 * j], yj = xy[2 * j + 1];
        const double xk = xy[2 * k], yk = xy[2 * k + 1];
        const double dij = dev_pdist(xi, yi, xj, yj), djk = dev_pdist(xj, yj, xk, yk), dik = dev_pdist(xi, yi, xk, yk);
        // stable ascending sort of (dij, djk, dik)  [std::stable_sort == insertion order on ties]
        double a = dij, b = djk, c = dik;
        if (b < a) { const double x = a; a = b; b = x; }
        if (c < b) {
            const double x = b; b = c; c = x;
            if (b < a) { const double y = a; a = b; b = y; }
        }
        if (!(a < min_side)) {
            m = b / a;
            l = c / a;
            // vertices ordered by the length of the opposite side, stable: (i, djk), (j, dik), (k, dij)
            int p0 = i, p1 = j, p2 = k;
            double q0 = djk, q1 = dik, q2 = dij;
            if (q1 < q0) { const double x = q0; q0 = q1; q1 = x; const int y = p0; p0 = p1; p1 = y; }
            if (q2 < q1) {
                const double x = q1; q1 = q2; q2 = x; const int y = p1; p1 = p2; p2 = y;
                if (q1 < q0) { const double x2 = q0; q0 = q1; q1 = x2; const int y2 = p0; p0 = p1; p1 = y2; }
            }
            v0 = (unsigned char)p0;
            v1 = (unsigned char)p1;
            v2 = (unsigned char)p2;
        }
    }
    mid[t] = m;
    lng[t] = l;
    srt[3 * t] = v0;
    srt[3 * t + 1] = v1;
    srt[3 * t + 2] = v2;
}

// Bucketed variant (the one match_stars launches).  Two triangles can only vote when their mid
// ratios differ by <= tol, so both lists are counting-sorted by bucket(mid) = floor(mid / (1.001 tol))
// (clamped to TRI_NBK - 1) and a reference triangle of bucket k meets only the target triangles of
// buckets k-1 .. k+1 -- one contiguous range of the sorted target list.  The exact f64 test of the
// brute-force kernel decides inside that range, and votes are sums, so the vote matrix is identical.
constexpr int TRI_NBK = 4096;
__device__ __forceinline__ int tri_bucket(double mid, double inv_w) {
    const double q = mid * inv_w;
    return q >= (double)(TRI_NBK - 1) ? TRI_NBK - 1 : (int)q;
}
__global__ void __launch_bounds__(256)
tri_hist_kernel(const double *__restrict__ r_mid, const double *__restrict__ t_mid, double inv_w,
                unsigned *__restrict__ hist) {  // hist[2][TRI_NBK]
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= TRI_MAX) return;
    const double rm = r_mid[t], tm = t_mid[t];
    if (fabs(rm) < CUDART_INF) atomicAdd(&hist[tri_bucket(rm, inv_w)], 1u);
    if (fabs(tm) < CUDART_INF) atomicAdd(&hist[TRI_NBK + tri_bucket(tm, inv_w)], 1u);
}
// block 0: reference list, block 1: target list.  start[side][TRI_NBK + 1], cursor[side][TRI_NBK]
__global__ void __launch_bounds__(1024)
tri_scan_kernel(const unsigned *__restrict__ hist, unsigned *__restrict__ start, unsigned *__restrict__ cursor) {
    __shared__ unsigned warp_sum[32];
    const unsigned *h = hist + blockIdx.x * TRI_NBK;
    unsigned *st = start + blockIdx.x * (TRI_NBK + 1), *cu = cursor + blockIdx.x * TRI_NBK;
    constexpr int PER = TRI_NBK / 1024;
    unsigned loc[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        loc[j] = h[threadIdx.x * PER + j];
        sum += loc[j];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned w = warp_sum[lane], winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += v;
        }
        warp_sum[lane] = winc - w;
    }
    __syncthreads();
    unsigned run = warp_sum[warp] + inc - sum;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        st[threadIdx.x * PER + j] = run;
        cu[threadIdx.x * PER + j] = run;
        run += loc[j];
    }
    if (threadIdx.x == 1023) st[TRI_NBK] = run;
}
struct TriList {
    double *mid, *lng;
    unsigned char *srt;
};
struct TriSorted {
    double2 *ml;  // (mid, long) per triangle
    unsigned char *srt;
};
__global__ void __launch_bounds__(256)
tri_scatter_kernel(TriList r_in, TriList t_in, double inv_w, unsigned *__restrict__ cursor, TriSorted r_out,
                   TriSorted t_out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= TRI_MAX) return;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const TriList &in = side ? t_in : r_in;
        const TriSorted &out = side ? t_out : r_out;
        const double m = in.mid[t];
        if (!(fabs(m) < CUDART_INF)) continue;
        const unsigned pos = atomicAdd(&cursor[side * TRI_NBK + tri_bucket(m, inv_w)], 1u);
        out.ml[pos] = make_double2(m, in.lng[t]);
        out.srt[3 * pos] = in.srt[3 * t];
        out.srt[3 * pos + 1] = in.srt[3 * t + 1];
        out.srt[3 * pos + 2] = in.srt[3 * t + 2];
    }
}
// One thread per bucket-sorted reference triangle; a warp walks the union of its lanes' target
// ranges, every lane testing the same (broadcast) target triangle.  The ranges of the densest
// buckets run to thousands of triangles, so blockIdx.y cuts the target list into TRI_SPAN-triangle
// spans and every warp takes the part of its range inside its span.
constexpr int TRI_SPAN = 512;
__global__ void __launch_bounds__(128)
tri_match_bucketed_kernel(TriSorted ref, TriSorted tgt, const unsigned *__restrict__ start, double inv_w,
                          double tol, unsigned *__restrict__ votes) {
    const unsigned n_ref = start[TRI_NBK];
    const unsigned *t_start = start + (TRI_NBK + 1);
    const unsigned a = blockIdx.x * blockDim.x + threadIdx.x;
    if ((a & ~31u) >= n_ref) return;  // whole warp beyond the list
    const bool live = a < n_ref;
    double rm = CUDART_INF, rl = CUDART_INF;  // (a dead lane never passes the test)
    unsigned lo = 0xffffffffu, hi = 0;
    if (live) {
        const double2 r = ref.ml[a];
        rm = r.x;
        rl = r.y;
        const int k = tri_bucket(rm, inv_w);
        lo = t_start[k > 0 ? k - 1 : 0];
        hi = t_start[k + 2 < TRI_NBK ? k + 2 : TRI_NBK];
    }
    const unsigned span0 = blockIdx.y * TRI_SPAN;
    const unsigned wlo = max(__reduce_min_sync(0xffffffffu, lo), span0);
    const unsigned whi = min(__reduce_max_sync(0xffffffffu, hi), span0 + TRI_SPAN);
    if (wlo >= whi) return;
    int r0 = 0, r1 = 0, r2 = 0;
    if (live) {
        r0 = ref.srt[3 * a];
        r1 = ref.srt[3 * a + 1];
        r2 = ref.srt[3 * a + 2];
    }
    auto test = [&](unsigned b, double2 t) {
        const double d_mid = fabs(rm - t.x);
        const double d_long = fabs(rl - t.y);
        if (d_mid > tol || d_long > tol) return;
        atomicAdd(&votes[r0 * MAX_STARS + tgt.srt[3 * b]], 1u);
        atomicAdd(&votes[r1 * MAX_STARS + tgt.srt[3 * b + 1]], 1u);
        atomicAdd(&votes[r2 * MAX_STARS + tgt.srt[3 * b + 2]], 1u);
    };
    unsigned b = wlo;
    for (; b + 4 <= whi; b += 4) {  // four broadcast loads in flight
        const double2 t0 = tgt.ml[b], t1 = tgt.ml[b + 1], t2 = tgt.ml[b + 2], t3 = tgt.ml[b + 3];
        test(b, t0);
        test(b + 1, t1);
        test(b + 2, t2);
        test(b + 3, t3);
    }
    for (; b < whi; ++b) test(b, tgt.ml[b]);
}

// match_triangles (affine.rs:320-384): triangles and votes on the GPU, greedy 1-1 assignment on the
// host.  Equal-vote pairs are taken in (ref index, tgt index) order (the reference's order is the
// HashMap iteration order, i.e. unspecified).
ab_status match_stars(ab_ctx *ctx, const std::vector<Pt> &rs, const std::vector<Pt> &ts,
                      cudaStream_t s, std::vector<Match> &matches) {
    matches.clear();
    if (rs.size() < 3 || ts.size() < 3) return AB_OK;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t b8 = al((size_t)TRI_MAX * 8), b3 = al((size_t)TRI_MAX * 3), bxy = al(2 * MAX_STARS * 8);
    const size_t o_rxy = 0, o_txy = bxy, o_rm = 2 * bxy, o_rl = o_rm + b8, o_rs = o_rl + b8, o_tm = o_rs + b3;
    const size_t o_tl = o_tm + b8, o_ts = o_tl + b8, o_v = o_ts + b3;
    // votes, then the bucket histograms (cleared with them), then starts / cursors / sorted lists
    const size_t o_h = o_v + al(MAX_STARS * MAX_STARS * 4), o_st = o_h + al(2 * TRI_NBK * 4);
    const size_t o_cu = o_st + al(2 * (TRI_NBK + 1) * 4), o_srm = o_cu + al(2 * TRI_NBK * 4);
    const size_t o_srs = o_srm + 2 * b8, o_stm = o_srs + b3, o_sts = o_stm + 2 * b8;  // (mid, long) pairs + vertices
    const size_t total = o_sts + b3;
    ab_status st = ab_scratch_reserve(ctx, 2, total);
    if (st != AB_OK) return st;
    char *b = static_cast<char *>(ctx->d_scratch[2]);
    static_assert(sizeof(Pt) == 16, "Pt is two doubles");
    AB_CUDA(ctx, cudaMemcpyAsync(b + o_rxy, rs.data(), rs.size() * sizeof(Pt), cudaMemcpyHostToDevice, s));
    AB_CUDA(ctx, cudaMemcpyAsync(b + o_txy, ts.data(), ts.size() * sizeof(Pt), cudaMemcpyHostToDevice, s));
    AB_CUDA(ctx, cudaMemsetAsync(b + o_v, 0, o_st - o_v, s));
    const bool trace = std::getenv("AB_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        cudaStreamSynchronize(s);
        fprintf(stderr, "[ab]   match: %s at %.3f ms\n", what,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    auto dp = [&](size_t off) { return reinterpret_cast<double *>(b + off); };
    auto bp = [&](size_t off) { return reinterpret_cast<unsigned char *>(b + off); };
    auto up = [&](size_t off) { return reinterpret_cast<unsigned *>(b + off); };
    const unsigned tb = (TRI_MAX + 127) / 128;
    build_tri_kernel<<<tb, 128, 0, s>>>(dp(o_rxy), (int)rs.size(), (double)INFINITY, MIN_TRIANGLE_SIDE, dp(o_rm),
                                       dp(o_rl), bp(o_rs));
    AB_LAUNCH_CHECK(ctx);
    build_tri_kernel<<<tb, 128, 0, s>>>(dp(o_txy), (int)ts.size(), -(double)INFINITY, MIN_TRIANGLE_SIDE, dp(o_tm),
                                       dp(o_tl), bp(o_ts));
    AB_LAUNCH_CHECK(ctx);
    lap("triangles built");
    if (std::getenv("AB_TRI_BRUTE")) {  // development switch: the all-pairs kernel (same votes)
        const dim3 grid((TRI_MAX + 255) / 256, (TRI_MAX + TRI_CHUNK - 1) / TRI_CHUNK);
        tri_match_kernel<<<grid, 256, 0, s>>>(dp(o_rm), dp(o_rl), bp(o_rs), TRI_MAX, dp(o_tm), dp(o_tl), bp(o_ts),
                                              TRI_MAX, TRIANGLE_TOLERANCE, up(o_v));
        AB_LAUNCH_CHECK(ctx);
    } else {
        const double inv_w = 1.0 / (TRIANGLE_TOLERANCE * 1.001);
        const TriList r_in{dp(o_rm), dp(o_rl), bp(o_rs)}, t_in{dp(o_tm), dp(o_tl), bp(o_ts)};
        const TriSorted r_out{reinterpret_cast<double2 *>(b + o_srm), bp(o_srs)};
        const TriSorted t_out{reinterpret_cast<double2 *>(b + o_stm), bp(o_sts)};
        tri_hist_kernel<<<(TRI_MAX + 255) / 256, 256, 0, s>>>(r_in.mid, t_in.mid, inv_w, up(o_h));
        AB_LAUNCH_CHECK(ctx);
        tri_scan_kernel<<<2, 1024, 0, s>>>(up(o_h), up(o_st), up(o_cu));
        AB_LAUNCH_CHECK(ctx);
        tri_scatter_kernel<<<(TRI_MAX + 255) / 256, 256, 0, s>>>(r_in, t_in, inv_w, up(o_cu), r_out, t_out);
        AB_LAUNCH_CHECK(ctx);
        const dim3 grid((TRI_MAX + 127) / 128, (TRI_MAX + TRI_SPAN - 1) / TRI_SPAN);
        tri_match_bucketed_kernel<<<grid, 128, 0, s>>>(r_out, t_out, up(o_st), inv_w, TRIANGLE_TOLERANCE, up(o_v));
        AB_LAUNCH_CHECK(ctx);
    }
    std::vector<unsigned> votes(MAX_STARS * MAX_STARS);
    lap("votes counted");
    AB_CUDA(ctx, cudaMemcpyAsync(votes.data(), b + o_v, votes.size() * 4, cudaMemcpyDeviceToHost, s));
    AB_CUDA(ctx, cudaStreamSynchronize(s));
    struct PV {
        int ri, ti;
        unsigned v;
    };
    std::vector<PV> pairs;
    for (int ri = 0; ri < (int)rs.size(); ++ri)
        for (int ti = 0; ti < (int)ts.size(); ++ti)
            if (votes[ri * MAX_STARS + ti]) pairs.push_back({ri, ti, votes[ri * MAX_STARS + ti]});
    std::stable_sort(pairs.begin(), pairs.end(), [](const PV &a, const PV &b) { return a.v > b.v; });
    std::vector<char> ur(rs.size(), 0), ut(ts.size(), 0);
    for (const PV &p : pairs) {
        if (ur[p.ri] || ut[p.ti]) continue;
        ur[p.ri] = 1;
        ut[p.ti] = 1;
        matches.push_back({rs[p.ri].x, rs[p.ri].y, ts[p.ti].x, ts[p.ti].y});
    }
    return AB_OK;
}

// the (up to) 120 brightest stars of a plane as the aligner sees it: normalize_for_detection fused
// into the detector's loads, detect_stars at 3.5 sigma, top MAX_STARS by flux (affine.rs:133-147)
ab_status align_star_points(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols, cudaStream_t s,
                            std::vector<Pt> &pts) {
    std::vector<ab_star> st;
    double bm, bs;
    Norm nm;
    const bool trace = std::getenv("AB_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    ab_status rc = normalize_params(ctx, d_img, rows * cols, s, &nm);
    if (rc != AB_OK) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    rc = detect_on_device(ctx, d_img, rows, cols, nm, DETECTION_SIGMA, s, st, &bm, &bs, MAX_STARS);
    if (rc != AB_OK) return rc;
    if (trace)
        fprintf(stderr, "[ab]   star points: normalise parameters %.3f ms, detect %.3f ms\n",
                std::chrono::duration<double, std::milli>(t1 - t0).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
    pts.clear();
    for (size_t i = 0; i < std::min(st.size(), (size_t)MAX_STARS); ++i) pts.push_back({st[i].x, st[i].y});
    return AB_OK;
}

// align_channel_affine (affine.rs:129-212) from both star lists onwards: triangle votes, RANSAC,
// sanity checks, phase-correlation fallback
ab_status align_from_star_points(ab_ctx *ctx, const std::vector<Pt> &rs, const std::vector<Pt> &ts,
                                 const float *d_ref, const float *d_tgt, size_t rows, size_t cols,
                                 int n_threads, cudaStream_t s, ab_align_result *out, double t_det) {
    const bool trace = std::getenv("AB_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(now() - t0).count();
    };
    if (rs.size() < MIN_MATCHES_RIGID || ts.size() < MIN_MATCHES_RIGID)
        return fallback(ctx, d_ref, d_tgt, rows, cols, s, out);
    const auto t1 = now();
    std::vector<Match> matches;
    ab_status st = match_stars(ctx, rs, ts, s, matches);
    if (st != AB_OK) return st;
    const double t_match = ms_since(t1);
    if (matches.size() < MIN_MATCHES_RIGID) return fallback(ctx, d_ref, d_tgt, rows, cols, s, out);
    const auto t2 = now();
    ab_align_result r;
    bool done = false;
    if (matches.size() >= MIN_MATCHES_AFFINE) {
        st = ransac_affine_device(ctx, matches, n_threads, s, &r, &done);
        if (st != AB_OK) return st;
        done = done && sane(r, rows, cols);
    }
    if (!done) done = ransac(matches, 1, n_threads, &r) && sane(r, rows, cols);
    if (trace)
        fprintf(stderr, "[ab] align %zux%zu: normalise+detect %.2f ms, triangles+votes+assign %.2f ms (%zu matches), "
                        "ransac %.2f ms\n", rows, cols, t_det, t_match, matches.size(), ms_since(t2));
    if (done) {
        *out = r;
        return AB_OK;
    }
    return fallback(ctx, d_ref, d_tgt, rows, cols, s, out);
}
// ... from the reference's star list onwards (the target's stars are detected here)
ab_status align_with_ref_points(ab_ctx *ctx, const std::vector<Pt> &rs, const float *d_ref,
                                const float *d_tgt, size_t rows, size_t cols, int n_threads,
                                cudaStream_t s, ab_align_result *out) {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<Pt> ts;
    ab_status st = align_star_points(ctx, d_tgt, rows, cols, s, ts);
    if (st != AB_OK) return st;
    const double t_det = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return align_from_star_points(ctx, rs, ts, d_ref, d_tgt, rows, cols, n_threads, s, out, t_det);
}

ab_status align_on_device(ab_ctx *ctx, const float *d_ref, const float *d_tgt, size_t rows,
                          size_t cols, int n_threads, cudaStream_t s, ab_align_result *out) {
    std::vector<Pt> rs;
    ab_status st = align_star_points(ctx, d_ref, rows, cols, s, rs);
    if (st != AB_OK) return st;
    return align_with_ref_points(ctx, rs, d_ref, d_tgt, rows, cols, n_threads, s, out);
}

ab_status upload(ab_ctx *ctx, int slot, const float *h, size_t len, float **d) {
    ab_status st = ab_scratch_reserve(ctx, slot, len * 4 + 256);
    if (st != AB_OK) return st;
    *d = static_cast<float *>(ctx->d_scratch[slot]);
    AB_CUDA(ctx, cudaMemcpyAsync(*d, h, len * 4, cudaMemcpyHostToDevice, ctx->stream));
    return AB_OK;
}

void fill_info(ab_detection_info *info, double bm, double bs, double thr, size_t rows, size_t cols,
               size_t n) {
    info->background_median = bm;
    info->background_sigma = bs;
    info->threshold_sigma = thr;
    info->image_width = cols;
    info->image_height = rows;
    info->n_stars = n;
}

}  // namespace

ab_status ab_detect_stars_on(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols,
                             double sigma_threshold, cudaStream_t s, std::vector<ab_star> &stars,
                             double *bg_median, double *bg_sigma) {
    Norm nm{0, 0.0, 1.0};
    return detect_on_device(ctx, d_img, rows, cols, nm, sigma_threshold, s, stars, bg_median, bg_sigma);
}

// ---- subframe scoring (core/analysis/subframe.rs:52-171): the detector + host scalar medians
namespace {
// median_of, subframe.rs:160-171: finite values, stable sort, even count -> mean of the middle two
double subframe_median(const std::vector<ab_star> &stars, double ab_star::*field) {
    std::vector<double> vals;
    for (const ab_star &s : stars)
        if (std::isfinite(s.*field)) vals.push_back(s.*field);
    if (vals.empty()) return 0.0;
    std::stable_sort(vals.begin(), vals.end());
    const size_t mid = vals.size() / 2;
    return vals.size() % 2 == 0 ? (vals[mid - 1] + vals[mid]) / 2.0 : vals[mid];
}
// compute_weight, subframe.rs:120-145
double subframe_weight(double fwhm, double ecc, double snr, double noise, const ab_subframe_config &c) {
    const double fwhm_score = fwhm > 0.5 ? 1.0 / fwhm : 0.0;
    const double ecc_score = 1.0 - ecc;
    const double snr_score = std::fmax(std::log(snr), 0.0);  // f64::max ignores NaN
    const double noise_score = 1.0 / (1.0 + noise * 10.0);
    const double total = c.fwhm_weight + c.eccentricity_weight + c.snr_weight + c.noise_weight;
    if (total < 1e-15) return 0.0;
    const double raw = c.fwhm_weight * fwhm_score + c.eccentricity_weight * ecc_score +
                       c.snr_weight * snr_score + c.noise_weight * noise_score;
    return std::fmax(raw / total, 0.0);
}
ab_status analyze_subframe_on(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols,
                              const ab_subframe_config &cfg, cudaStream_t s, ab_subframe_metrics *out) {
    std::vector<ab_star> stars;
    double bm = 0.0, bs = 1.0;
    Norm nm{0, 0.0, 1.0};
    ab_status st = detect_on_device(ctx, d_img, rows, cols, nm, 4.0, s, stars, &bm, &bs);  // DETECTION_SIGMA :6
    if (st != AB_OK) return st;
    *out = {};
    out->star_count = stars.size();
    out->background_median = bm;
    out->background_sigma = bs;
    const size_t min_needed = std::min<size_t>(5, (size_t)cfg.min_stars);  // MIN_STARS_FOR_METRICS :7
    if (stars.size() < min_needed) return AB_OK;
    out->median_fwhm = subframe_median(stars, &ab_star::fwhm);
    out->median_eccentricity = subframe_median(stars, &ab_star::eccentricity);
    out->median_snr = subframe_median(stars, &ab_star::snr);
    out->noise_ratio = bm > 1e-15 ? bs / bm : 0.0;
    out->weight = subframe_weight(out->median_fwhm, out->median_eccentricity, out->median_snr,
                                  out->noise_ratio, cfg);
    out->accepted = (stars.size() >= (size_t)cfg.min_stars && out->median_fwhm <= cfg.max_fwhm &&
                     out->median_eccentricity <= cfg.max_eccentricity && out->median_snr >= cfg.min_snr)
                        ? 1
                        : 0;
    return AB_OK;
}
}  // namespace

extern "C" {
ab_status ab_analyze_subframe_dev(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols,
                                  const ab_subframe_config *cfg, ab_subframe_metrics *out, void *stream) {
    if (!ctx || !cfg || !out || ((rows * cols) && !d_img)) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    return analyze_subframe_on(ctx, d_img, rows, cols, *cfg, (cudaStream_t)stream, out);
}
ab_status ab_analyze_subframe(ab_ctx *ctx, const float *img, size_t rows, size_t cols,
                              const ab_subframe_config *cfg, ab_subframe_metrics *out) {
    if (!ctx || !cfg || !out || ((rows * cols) && !img)) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    float *d = nullptr;
    if (rows * cols) {
        ab_status st = upload(ctx, 0, img, rows * cols, &d);
        if (st != AB_OK) return st;
    }
    return analyze_subframe_on(ctx, d, rows, cols, *cfg, ctx->stream, out);
}
void ab_normalize_weights(ab_subframe_metrics *metrics, size_t n) {
    if (!metrics) return;
    double max_w = 0.0;
    for (size_t i = 0; i < n; ++i) max_w = std::fmax(max_w, metrics[i].weight);
    if (max_w > 1e-15)
        for (size_t i = 0; i < n; ++i) metrics[i].weight /= max_w;
}
}  // extern "C"

extern "C" {

ab_status ab_detect_stars_dev(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols,
                              double sigma_threshold, ab_star *stars, size_t max_stars,
                              ab_detection_info *info, void *stream) {
    if (!ctx || !info || (max_stars && !stars) || (rows * cols && !d_img)) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    std::vector<ab_star> found;
    double bm, bs;
    Norm nm{0, 0.0, 1.0};
    ab_status st = detect_on_device(ctx, d_img, rows, cols, nm, sigma_threshold,
                                    (cudaStream_t)stream, found, &bm, &bs);
    if (st != AB_OK) return st;
    for (size_t i = 0; i < std::min(found.size(), max_stars); ++i) stars[i] = found[i];
    fill_info(info, bm, bs, sigma_threshold, rows, cols, found.size());
    return AB_OK;
}

ab_status ab_detect_stars(ab_ctx *ctx, const float *img, size_t rows, size_t cols,
                          double sigma_threshold, ab_star *stars, size_t max_stars,
                          ab_detection_info *info) {
    if (!ctx || !info || (max_stars && !stars) || (rows * cols && !img)) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    std::vector<ab_star> found;
    double bm = 0.0, bs = 1.0;
    if (rows >= 3 && cols >= 3) {
        float *d = nullptr;
        ab_status st = upload(ctx, 0, img, rows * cols, &d);
        if (st != AB_OK) return st;
        Norm nm{0, 0.0, 1.0};
        st = detect_on_device(ctx, d, rows, cols, nm, sigma_threshold, ctx->stream, found, &bm, &bs);
        if (st != AB_OK) return st;
    }
    for (size_t i = 0; i < std::min(found.size(), max_stars); ++i) stars[i] = found[i];
    fill_info(info, bm, bs, sigma_threshold, rows, cols, found.size());
    return AB_OK;
}

// estimate_background, star_detection.rs:32-84, for a caller-chosen tile size
ab_status ab_estimate_background(ab_ctx *ctx, const float *img, size_t rows, size_t cols,
                                 size_t tile_size, double *median, double *sigma) {
    if (!ctx || !img || !median || !sigma) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    *median = 0.0;
    *sigma = 1.0;
    if (rows == 0 || cols == 0) return AB_OK;
    float *d = nullptr;
    ab_status st = upload(ctx, 0, img, rows * cols, &d);
    if (st != AB_OK) return st;
    const size_t step = std::max(tile_size, (size_t)16);
    const size_t tiles_y = (rows + step - 1) / step, tiles_x = (cols + step - 1) / step;
    const size_t n_tiles = tiles_y * tiles_x;
    st = ab_scratch_reserve(ctx, 1, n_tiles * 20 + 256);
    if (st != AB_OK) return st;
    char *base = static_cast<char *>(ctx->d_scratch[1]);
    double *d_med = reinterpret_cast<double *>(base), *d_sig = reinterpret_cast<double *>(base + n_tiles * 8);
    int *d_ok = reinterpret_cast<int *>(base + n_tiles * 16);
    Norm nm{0, 0.0, 1.0};
    st = launch_bg_tiles(ctx, d, rows, cols, step, tiles_x, n_tiles, nm, d_med, d_sig, d_ok, ctx->stream);
    if (st != AB_OK) return st;
    std::vector<double> h_med(n_tiles), h_sig(n_tiles);
    std::vector<int> h_ok(n_tiles);
    AB_CUDA(ctx, cudaMemcpyAsync(h_med.data(), d_med, n_tiles * 8, cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaMemcpyAsync(h_sig.data(), d_sig, n_tiles * 8, cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaMemcpyAsync(h_ok.data(), d_ok, n_tiles * 4, cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<double> medians, sigmas;
    for (size_t t = 0; t < n_tiles; ++t)
        if (h_ok[t]) {
            medians.push_back(h_med[t]);
            sigmas.push_back(h_sig[t]);
        }
    if (medians.empty()) return AB_OK;
    std::sort(medians.begin(), medians.end());
    std::sort(sigmas.begin(), sigmas.end());
    *median = medians[medians.size() / 2];
    *sigma = std::fmax(sigmas[sigmas.size() / 2], 1e-10);
    return AB_OK;
}

// normalize_for_detection, affine.rs:24-53 (materialised; the aligner fuses it into its loads)
ab_status ab_normalize_for_detection(ab_ctx *ctx, const float *img, size_t len, float *out) {
    if (!ctx || (len && (!img || !out))) return AB_ERR_INVALID;
    if (len == 0) return AB_OK;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    float *d = nullptr;
    ab_status st = upload(ctx, 0, img, len, &d);
    if (st != AB_OK) return st;
    Norm nm;
    st = normalize_params(ctx, d, len, ctx->stream, &nm);
    if (st != AB_OK) return st;
    if (!nm.on) {
        std::memcpy(out, img, len * 4);
        return AB_OK;
    }
    st = ab_scratch_reserve(ctx, 1, len * 4 + 256);
    if (st != AB_OK) return st;
    float *d_out = static_cast<float *>(ctx->d_scratch[1]);
    normalize_kernel<<<(unsigned)ctx->sm_count * 8, 256, 0, ctx->stream>>>(d, len, nm, d_out);
    AB_LAUNCH_CHECK(ctx);
    AB_CUDA(ctx, cudaMemcpyAsync(out, d_out, len * 4, cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}

ab_status ab_align_affine_dev(ab_ctx *ctx, const float *d_ref, const float *d_tgt, size_t rows,
                              size_t cols, int n_threads, ab_align_result *out, void *stream) {
    if (!ctx || !out || !d_ref || !d_tgt) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    return align_on_device(ctx, d_ref, d_tgt, rows, cols, n_threads, (cudaStream_t)stream, out);
}

// The reference plane's star list once (align_channels_cmd aligns every channel against the same
// reference, cmd/compose/blend.rs:263-296; the reference re-detects it per pair, :279): xy_out =
// up to 120 (x, y) pairs, *n_out = stars written.  ab_align_affine_with_ref_dev then equals
// ab_align_affine_dev(ref, tgt) bit for bit.  A frames-across-GPUs run broadcasts these <= 1920
// bytes instead of the reference frame.
ab_status ab_align_ref_stars_dev(ab_ctx *ctx, const float *d_ref, size_t rows, size_t cols,
                                 double *xy_out, size_t *n_out, void *stream) {
    if (!ctx || !d_ref || !xy_out || !n_out) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    std::vector<Pt> pts;
    ab_status st = align_star_points(ctx, d_ref, rows, cols, (cudaStream_t)stream, pts);
    if (st != AB_OK) return st;
    for (size_t i = 0; i < pts.size(); ++i) {
        xy_out[2 * i] = pts[i].x;
        xy_out[2 * i + 1] = pts[i].y;
    }
    *n_out = pts.size();
    return AB_OK;
}
// d_ref may be NULL when the reference plane is not on this GPU: the phase-correlation fallback
// (affine.rs:243-270) then cannot run and the result falls back to identity (method 3).
ab_status ab_align_affine_with_ref_dev(ab_ctx *ctx, const double *ref_xy, size_t n_ref,
                                       const float *d_ref, const float *d_tgt, size_t rows, size_t cols,
                                       int n_threads, ab_align_result *out, void *stream) {
    if (!ctx || !out || !d_tgt || (n_ref && !ref_xy)) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    std::vector<Pt> rs;
    for (size_t i = 0; i < std::min(n_ref, (size_t)MAX_STARS); ++i) rs.push_back({ref_xy[2 * i], ref_xy[2 * i + 1]});
    return align_with_ref_points(ctx, rs, d_ref, d_tgt, rows, cols, n_threads, (cudaStream_t)stream, out);
}

// Both star lists given (each from ab_align_ref_stars_dev on its plane): a rank can detect its
// target frames before the reference's list has arrived.  d_ref / d_tgt feed only the
// phase-correlation fallback (either NULL: identity, method 3).  Equals ab_align_affine_dev.
ab_status ab_align_affine_from_stars_dev(ab_ctx *ctx, const double *ref_xy, size_t n_ref, const double *tgt_xy,
                                         size_t n_tgt, const float *d_ref, const float *d_tgt, size_t rows,
                                         size_t cols, int n_threads, ab_align_result *out, void *stream) {
    if (!ctx || !out || (n_ref && !ref_xy) || (n_tgt && !tgt_xy)) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    std::vector<Pt> rs, ts;
    for (size_t i = 0; i < std::min(n_ref, (size_t)MAX_STARS); ++i) rs.push_back({ref_xy[2 * i], ref_xy[2 * i + 1]});
    for (size_t i = 0; i < std::min(n_tgt, (size_t)MAX_STARS); ++i) ts.push_back({tgt_xy[2 * i], tgt_xy[2 * i + 1]});
    return align_from_star_points(ctx, rs, ts, d_tgt ? d_ref : nullptr, d_tgt, rows, cols, n_threads,
                                  (cudaStream_t)stream, out, 0.0);
}

ab_status ab_align_affine(ab_ctx *ctx, const float *ref, const float *tgt, size_t rows, size_t cols,
                          int n_threads, ab_align_result *out) {
    if (!ctx || !out || !ref || !tgt) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    // both planes side by side in scratch slot 0
    const size_t plane = (rows * cols * 4 + 255) / 256 * 256;
    ab_status st = ab_scratch_reserve(ctx, 0, plane * 2 + 256);
    if (st != AB_OK) return st;
    float *d_ref = static_cast<float *>(ctx->d_scratch[0]);
    float *d_tgt = reinterpret_cast<float *>(static_cast<char *>(ctx->d_scratch[0]) + plane);
    AB_CUDA(ctx, cudaMemcpyAsync(d_ref, ref, rows * cols * 4, cudaMemcpyHostToDevice, ctx->stream));
    AB_CUDA(ctx, cudaMemcpyAsync(d_tgt, tgt, rows * cols * 4, cudaMemcpyHostToDevice, ctx->stream));
    return align_on_device(ctx, d_ref, d_tgt, rows, cols, n_threads, ctx->stream, out);
}

// match_triangles on caller-provided star lists (x, y pairs): matches_out = 4 doubles per match
ab_status ab_match_stars(ab_ctx *ctx, const double *ref_xy, size_t n_ref, const double *tgt_xy,
                         size_t n_tgt, double *matches_out, size_t max_matches, size_t *n_matches) {
    if (!ctx || !n_matches || (n_ref && !ref_xy) || (n_tgt && !tgt_xy)) return AB_ERR_INVALID;
    if (n_ref > MAX_STARS || n_tgt > MAX_STARS)
        return ab_fail(ctx, AB_ERR_INVALID, "at most %d stars per list (MAX_STARS)", MAX_STARS);
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    std::vector<Pt> rs(n_ref), ts(n_tgt);
    for (size_t i = 0; i < n_ref; ++i) rs[i] = {ref_xy[2 * i], ref_xy[2 * i + 1]};
    for (size_t i = 0; i < n_tgt; ++i) ts[i] = {tgt_xy[2 * i], tgt_xy[2 * i + 1]};
    std::vector<Match> m;
    ab_status st = match_stars(ctx, rs, ts, ctx->stream, m);
    if (st != AB_OK) return st;
    *n_matches = m.size();
    for (size_t i = 0; i < std::min(m.size(), max_matches); ++i) {
        matches_out[4 * i] = m[i].rx;
        matches_out[4 * i + 1] = m[i].ry;
        matches_out[4 * i + 2] = m[i].tx;
        matches_out[4 * i + 3] = m[i].ty;
    }
    return AB_OK;
}

// ransac_affine (affine.rs:400-517), host scalar code: method 0 affine / 1 rigid; returns
// AB_OK with *found = 0 when the reference returns None
ab_status ab_ransac_affine(const double *matches, size_t n, int method, int n_threads,
                           ab_align_result *out, int *found) {
    if (!out || !found || (n && !matches)) return AB_ERR_INVALID;
    std::vector<Match> m(n);
    for (size_t i = 0; i < n; ++i)
        m[i] = {matches[4 * i], matches[4 * i + 1], matches[4 * i + 2], matches[4 * i + 3]};
    *found = ransac(m, method, n_threads, out) ? 1 : 0;
    return AB_OK;
}
// the same result with the hypotheses of the affine model evaluated on the GPU (what ab_align_affine
// runs); the rigid model is host code in both
ab_status ab_ransac_affine_gpu(ab_ctx *ctx, const double *matches, size_t n, int method, int n_threads,
                               ab_align_result *out, int *found) {
    if (!ctx || !out || !found || (n && !matches)) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    std::vector<Match> m(n);
    for (size_t i = 0; i < n; ++i)
        m[i] = {matches[4 * i], matches[4 * i + 1], matches[4 * i + 2], matches[4 * i + 3]};
    if (method != 0) {
        *found = ransac(m, method, n_threads, out) ? 1 : 0;
        return AB_OK;
    }
    bool f = false;
    AB_CUDA(ctx, cudaSetDevice(ctx->device));
    const ab_status st = ransac_affine_device(ctx, m, n_threads, ctx->stream, out, &f);
    *found = f ? 1 : 0;
    return st;
}

}  // extern "C"
