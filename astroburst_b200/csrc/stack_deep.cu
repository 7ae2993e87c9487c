// stack_deep.cu -- sigma-clipped stack for frame counts without a register kernel (33..256 frames,
// BASELINE config 5: 256 x 4096^2) and the per-pixel median combine of the calibration masters.
//
// Reference semantics: stack_images / sigma_clip_combine,
// src-tauri/src/core/stacking/combine.rs:14-92,160-182 (align=false core loop);
// median_combine_row_major, src-tauri/src/core/stacking/calibration.rs:85-123.
//
// Layout of the work (R = samples per lane: 8/16/32/64 for <= 32/64/128/256 frames):
//   * FOUR lanes cooperate on one pixel, a warp owns 8 pixels, a 4-warp block owns a 32-pixel tile
//     (one 128-byte line of every frame).  The block streams tiles [4R frames x 32 px] into a
//     double-buffered shared-memory stage with 16-byte cp.async (LDGSTS): every warp instruction
//     fetches four whole 128-byte lines, so the same code runs at full efficiency on PEER pointers
//     (frame-sharded stacks read the owners' HBM over NVLink, no staging copy); the fetch of tile
//     i+1 overlaps all the arithmetic of tile i, one __syncthreads per tile.
//   * each lane copies its R samples to registers (conflict-free LDS), non-finite -> +inf, and
//     sorts them with the register network of stack_kernels.cuh; two bitonic merge levels across
//     the lane quad (flip stage = 2 SHFL + 2 predicated FMNMX per register pair, then in-lane
//     half-cleaners) leave the pixel's 4R samples fully sorted: rank g at lane g / R, register
//     g % R.  The sorted run is also written to the warp's own (consumed) part of the stage for
//     the few dynamically indexed reads below.
//   * iteration 0 (combine.rs:37-48): upper median S[n/2]; the upper-median MAD is
//     min_t max(med - S[t], S[t + n/2] - med) and, the first term falling and the second rising
//     with t, the minimum sits at their crossing: a binary search instead of a scan.
//   * the clip (combine.rs:65-82): dev = v - center is monotone in v, so the rejects are a prefix
//     and a suffix of the surviving rank range [a, b); the quad probes four candidates per step.
//   * iterations >= 1 (combine.rs:49-61): moments D1 = sum (v - c0), D2 = sum (v - c0)^2 about
//     c0 = median, accumulated in f64 from the registers over the survivors of ITERATION 0 (so no
//     far outlier ever enters them) and downdated by the few samples later iterations remove;
//     mean = (len c0 + D1) / len exactly (the reference's f64 sum is exact for f32 samples of
//     comparable magnitude), sum (v - mean)^2 = D2 - 2 d D1 + len d^2 with d = mean - c0.  A
//     downdate that removes most of D2 (cancellation) triggers a fresh accumulation instead.
#include "stack_kernels.cuh"

#ifndef AB_DEEP_MN
#define AB_DEEP_MN 2
#define AB_DEEP_MD 3
#endif
#ifndef AB_DEEP_XIMAD  // cross-lane stages: upper lanes take a + b - min on the FMA pipe (predicated IMADs)
#define AB_DEEP_XIMAD 0
#endif
#ifndef AB_DEEP_MAD4   // MAD crossing search: 1 = quad-cooperative 5-ary search, 0 = binary search
#define AB_DEEP_MAD4 1
#endif

namespace {

using abk::CeConst;

constexpr int DL = 4;        // lanes per pixel
constexpr int DPX = 32 / DL; // pixels per warp
#ifndef AB_DEEP_WARPS
#define AB_DEEP_WARPS 4
#endif
constexpr int DWARPS = AB_DEEP_WARPS;  // warps per block
constexpr int DTHREADS = DWARPS * 32;
constexpr int TILE_PX = DWARPS * DPX;  // 32 pixels = one 128-byte line per frame
constexpr int ROW_CHUNKS = TILE_PX / 4;  // 16-byte chunks of one frame row of a tile
constexpr unsigned FULL = 0xffffffffu;

template <int R>
struct Geom {
    static constexpr int LOG2R = R == 64 ? 6 : (R == 32 ? 5 : (R == 16 ? 4 : 3));
    static constexpr int FR = DL * R;             // frame rows of a tile (frame count padded)
    static constexpr int LANE_PITCH = R + 4;      // sorted layout: words per lane block
    static constexpr int PITCH = DL * R + 16;     // sorted layout: words per pixel
    // words per warp region: the tile part [FR][8 px] (32 R words) overlaid by the sorted layout
    // (8 PITCH = 32 R + 128 words); + 8 so the four regions start 8 banks apart (cp.async writes)
    static constexpr int REGION = DPX * PITCH + 8;
    static constexpr int STAGE = DWARPS * REGION;
    static constexpr size_t SMEM = (size_t)2 * STAGE * sizeof(float) + AB_MAX_FRAMES * sizeof(float *);
};

// 16-byte global -> shared copies.  Local frames bypass L1 (.cg); frames that live in a PEER's HBM
// allocate in L1 (.ca): on peer addresses the L1 turns a warp's 16-byte pieces into whole-line
// NVLink requests (measured at N = 2: 244 GB/s with .cg, 412 GB/s with .ca; local cost +2 %).
__device__ __forceinline__ void cp_async16(float *dst, const float *src, bool peer) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    if (peer) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(float *dst, const float *src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Block-cooperative fetch of tile `tile` (pixels [32 tile, 32 tile + 32) of every frame) into
// `stage`: frame f, 16-byte chunk k (pixels 4k..4k+3) -> region k/2, word f*8 + (k&1)*4.
// Out of line (called when no sample registers are live): one copy of the address arithmetic.
template <int R>
static __device__ __noinline__ void fill_stage(float *stage, const float *const *ptrs, int n_frames,
                                               size_t npix, size_t tile, bool aligned16, bool peer) {
    using G = Geom<R>;
    const size_t px0 = tile * TILE_PX;
    constexpr int ITERS = (G::FR * ROW_CHUNKS) / DTHREADS;
    if (aligned16 && px0 + TILE_PX <= npix) {
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int c = it * DTHREADS + (int)threadIdx.x;
            const int f = c / ROW_CHUNKS, k = c % ROW_CHUNKS;
            float *dst = stage + (k >> 1) * G::REGION + f * 8 + (k & 1) * 4;
            if (f < n_frames) cp_async16(dst, ptrs[f] + px0 + 4 * k, peer);
            else  // rows past the frame count: +inf sorts last and counts as "not finite"
                *reinterpret_cast<float4 *>(dst) =
                    make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
        }
    } else {  // unaligned frames or the ragged last tile: 4-byte copies, pixels past the end untouched
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
            const int c = it * DTHREADS + (int)threadIdx.x;
            const int f = c / ROW_CHUNKS, k = c % ROW_CHUNKS;
            float *dst = stage + (k >> 1) * G::REGION + f * 8 + (k & 1) * 4;
            const size_t px = px0 + 4 * k;
#pragma unroll 1
            for (int e = 0; e < 4; ++e) {
                if (f >= n_frames) dst[e] = CUDART_INF_F;
                else if (px + e < npix) cp_async4(dst + e, ptrs[f] + px + e);
            }
        }
    }
    cp_async_commit();
}

// min(own, x) on the lower lane, max(own, x) on the upper lane.  Operands are finite or +inf.
__device__ __forceinline__ float keep_side(float own, float x, bool upper, const CeConst &kc) {
#if AB_DEEP_XIMAD
    float m = fminf(own, x);
    if (upper) {  // own + x - min on the integer bit patterns: exactly the other operand
        const int t = __float_as_int(own) * kc.one + __float_as_int(x);
        m = __int_as_float(__float_as_int(m) * kc.mone + t);
    }
    return m;
#else
    return upper ? fmaxf(own, x) : fminf(own, x);
#endif
}
// One bitonic "flip" stage across lanes (lane ^ xm): register i meets the partner's register
// R-1-i; the lower lane keeps the minima, the upper lane the maxima.
template <int R>
__device__ __forceinline__ void flip_stage(float (&s)[R], int xm, bool upper, const CeConst &kc) {
#pragma unroll
    for (int i = 0; i < R / 2; ++i) {
        const float x = __shfl_xor_sync(FULL, s[i], xm);
        const float y = __shfl_xor_sync(FULL, s[R - 1 - i], xm);
        s[i] = keep_side(s[i], y, upper, kc);
        s[R - 1 - i] = keep_side(s[R - 1 - i], x, upper, kc);
    }
}
// Half-cleaner across lanes (stride = R elements): register i meets the partner's register i.
template <int R>
__device__ __forceinline__ void straight_stage(float (&s)[R], int xm, bool upper, const CeConst &kc) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const float x = __shfl_xor_sync(FULL, s[i], xm);
        s[i] = keep_side(s[i], x, upper, kc);
    }
}
// In-lane half-cleaners, strides R/2 .. 1: a bitonic run of R registers -> ascending.
template <int R, int STRIDE, int... I>
__device__ __forceinline__ void halfclean_stride(float (&s)[R], const CeConst &kc,
                                                 std::integer_sequence<int, I...>) {
    // pair I (0 .. R/2-1) of this stride: low index = (I / STRIDE) * 2 STRIDE + I % STRIDE
    (abk::ce<abk::ce_is_imad<AB_DEEP_MN, AB_DEEP_MD>(I)>(
         s[(I / STRIDE) * 2 * STRIDE + (I % STRIDE)],
         s[(I / STRIDE) * 2 * STRIDE + (I % STRIDE) + STRIDE], kc),
     ...);
}
template <int R, int STRIDE>
__device__ __forceinline__ void halfclean(float (&s)[R], const CeConst &kc) {
    if constexpr (STRIDE >= 1) {
        halfclean_stride<R, STRIDE>(s, kc, std::make_integer_sequence<int, R / 2>{});
        halfclean<R, STRIDE / 2>(s, kc);
    }
}

__device__ __forceinline__ double quad_sum(double v) {
    v += __shfl_xor_sync(FULL, v, 1);
    v += __shfl_xor_sync(FULL, v, 2);
    return v;
}

// PRE = false is the plain stack (no calibration terms, no median-only mode compiled in).
template <int R, bool PRE>
__global__ void __launch_bounds__(DTHREADS, (R == 64 ? 12 : 16) / DWARPS)
stack_deep_kernel(const __grid_constant__ FrameTable ft, int n_frames, size_t npix,
                  float sigma_low, float sigma_high, unsigned max_iter, CeConst kc,
                  const __grid_constant__ DeepPre pre, int aligned16, float *__restrict__ out,
                  unsigned long long *__restrict__ rejected_total) {
    using G = Geom<R>;
    extern __shared__ __align__(16) float deep_smem[];
    const float **ptrs = reinterpret_cast<const float **>(deep_smem + 2 * G::STAGE);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int p = lane >> 2;  // pixel of the warp
    const int l = lane & 3;   // lane of the pixel's quad
    const bool has_bias = PRE && pre.bias != nullptr;
    const bool has_dark = PRE && pre.dark != nullptr;
    const bool median_only = PRE && pre.median_only != 0;
    unsigned local_rej = 0;

    for (int f = threadIdx.x; f < AB_MAX_FRAMES; f += DTHREADS) ptrs[f] = ft.p[f];
    __syncthreads();

    const size_t n_tiles = (npix + TILE_PX - 1) / TILE_PX;
    size_t tile = blockIdx.x;
    const bool al16 = (aligned16 & 1) != 0, peer = (aligned16 & 2) != 0;  // bit 1: some frames are peer memory
    if (tile < n_tiles) fill_stage<R>(deep_smem, ptrs, n_frames, npix, tile, al16, peer);
    int buf = 0;
    for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        cp_async_wait_all();
        __syncthreads();  // tile `tile` has landed; every warp is done with the other stage
        {  // the next tile streams in while this one is worked on (no sample registers live yet)
            const size_t next = tile + gridDim.x;
            if (next < n_tiles)
                fill_stage<R>(deep_smem + (buf ^ 1) * G::STAGE, ptrs, n_frames, npix, next, al16, peer);
        }
        float *region = deep_smem + buf * G::STAGE + warp * G::REGION;
        const size_t pix = tile * TILE_PX + warp * DPX + p;
        const bool in_image = pix < npix;

        // ---------------------------------------------------------------- samples -> registers
        // lane l takes frames [R l, R l + R), rotated by l rows so that the quad's four loads of
        // a step fall in different banks (word = 8 f + p)
        float s[R];
        {
            const float *src = region + (R * l + l) * 8 + p;
#pragma unroll
            for (int j = 0; j < R; ++j) {
                if (j < R - 3) s[j] = src[j * 8];
                else s[j] = src[(j + l >= R ? j - R : j) * 8];
            }
        }
        if constexpr (PRE) {  // calibration.rs:160-162,199-206: frame - bias, then - dark * ratio
            const float pre_b = (in_image && has_bias) ? pre.bias[pix] : 0.0f;
            const float pre_d = (in_image && has_dark) ? pre.dark[pix] * pre.dark_ratio : 0.0f;
            if (has_bias) {
#pragma unroll
                for (int j = 0; j < R; ++j) s[j] -= pre_b;
            }
            if (has_dark) {
#pragma unroll
                for (int j = 0; j < R; ++j) s[j] -= pre_d;
            }
        }
        __syncwarp();  // the whole warp has read its part of the stage: it is scratch from here on

        // non-finite samples (combine.rs:170-175 drops them) -> +inf; n = finite samples of the pixel
        int n;
        {
            float probe = 0.0f;
#pragma unroll
            for (int j = 0; j < R; ++j) probe = __fmaf_rn(s[j], 0.0f, probe);  // NaN iff any non-finite
            const bool dirty = n_frames < G::FR || __any_sync(FULL, !(probe == 0.0f));
            if (dirty) {
                int nl = 0;
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const bool fin = fabsf(s[j]) < CUDART_INF_F;
                    nl += fin ? 1 : 0;
                    s[j] = fin ? s[j] : CUDART_INF_F;
                }
                nl += __shfl_xor_sync(FULL, nl, 1);
                nl += __shfl_xor_sync(FULL, nl, 2);
                n = nl;
            } else {
                n = G::FR;
            }
        }

        // ---------------------------------------------------------------- sort + merge
        abk::sort_network<R, AB_DEEP_MN, AB_DEEP_MD>(s, kc);
#pragma unroll 1
        for (int level = 0; level < 2; ++level) {
            flip_stage<R>(s, level ? 3 : 1, (lane & (level ? 2 : 1)) != 0, kc);
            if (level) straight_stage<R>(s, 1, (lane & 1) != 0, kc);
            halfclean<R, R / 2>(s, kc);
        }
        // rank g of the pixel is now register g % R of lane g / R; mirror it in shared memory
        float *srt = region + p * G::PITCH;
        {
            float *dst = srt + l * G::LANE_PITCH;
#pragma unroll
            for (int j = 0; j < R; j += 4)
                *reinterpret_cast<float4 *>(dst + j) = make_float4(s[j], s[j + 1], s[j + 2], s[j + 3]);
        }
        __syncwarp();
        auto S = [&](int g) -> float { return srt[g + 4 * (g >> G::LOG2R)]; };

        // ---------------------------------------------------------------- iteration 0
        const int h = n >> 1;
        const float med = n > 0 ? S(h) : 0.0f;  // upper median (combine.rs:38-40); n == 1: the sample
        float result = med;                      // n == 0 -> 0, n == 1 -> the sample (combine.rs:21-26)
        unsigned rejected = 0;
        if (!median_only) {
            int a = 0, b = n;
            bool active = n >= 2 && max_iter > 0;
            float sigma = 0.0f, last_center = 0.0f;
            if (__any_sync(FULL, active)) {
                // MAD = min_t max(med - S[t], S[t+h] - med), t in [0, n-h): first t whose right
                // term reaches the left one, then the better of t and t-1
                const int W = n - h;
#if AB_DEEP_MAD4
                // quad-cooperative search for t* = first t in [0, W) with right(t) >= left(t) (W if
                // none): the lanes probe the last element of four consecutive blocks of `st`
                // candidates; the count of leading "false" picks the block
                int lo_t = 0;
#pragma unroll 1
                for (int st = G::FR / 8;; st = st > 2 ? st >> 2 : 1) {
                    const int q = lo_t + (l + 1) * st - 1;
                    const bool below = active && q < W && !(S(q + h) - med >= med - S(q));
                    const unsigned bq = (__ballot_sync(FULL, below) >> (4 * p)) & 0xfu;
                    lo_t += __popc(bq) * st;
                    if (st == 1) break;
                }
                lo_t = min(lo_t, W);
#else
                int lo_t = 0, hi_t = active ? W : 0;
#pragma unroll 1
                for (int r = 0; r < G::LOG2R + 2; ++r) {
                    const int mid = (lo_t + hi_t) >> 1;
                    if (lo_t < hi_t) {
                        if (S(mid + h) - med >= med - S(mid)) hi_t = mid;
                        else lo_t = mid + 1;
                    }
                }
#endif
                float mad = CUDART_INF_F;
                if (active && lo_t < W) mad = fmaxf(med - S(lo_t), S(lo_t + h) - med);
                if (active && lo_t > 0)
                    mad = fminf(mad, fmaxf(med - S(lo_t - 1), S(lo_t - 1 + h) - med));
                sigma = fmaxf((float)((double)mad * abk::MAD_TO_SIGMA), 1e-10f);
            }
            const double c0 = (double)med;
            double D1 = 0.0, D2 = 0.0, D2_fresh = 0.0;
            float center = med;
            for (unsigned it = 0;; ++it) {
                // ---- one clip pass (combine.rs:62-82) for the quads still iterating
                int removed = 0;
                if (it < max_iter && __any_sync(FULL, active)) {
                    if (active && b - a < 2) active = false;  // combine.rs:33-35
                    if (active) last_center = center;
                    const float lo = -sigma_low * sigma;
                    const float hi = sigma_high * sigma;
                    // negated compares: NaN bounds reject everything, like `dev >= lo && dev <= hi`
                    int na = a, nb = b;
                    bool search = active;
#pragma unroll 1
                    while (__any_sync(FULL, search)) {
                        const int g = na + l;
                        const bool rj = search && g < nb && !(S(g) - center >= lo);
                        const unsigned q = (__ballot_sync(FULL, rj) >> (4 * p)) & 0xfu;
                        const int cnt = q == 0xfu ? 4 : __ffs(~q) - 1;
                        if (search) na += cnt;
                        search = search && cnt == 4 && na < nb;
                    }
                    search = active && na < nb;
#pragma unroll 1
                    while (__any_sync(FULL, search)) {
                        const int g = nb - 1 - l;
                        const bool rj = search && g >= na && !(S(g) - center <= hi);
                        const unsigned q = (__ballot_sync(FULL, rj) >> (4 * p)) & 0xfu;
                        const int cnt = q == 0xfu ? 4 : __ffs(~q) - 1;
                        if (search) nb -= cnt;
                        search = search && cnt == 4 && nb > na;
                    }
                    if (active) {
                        removed = (b - a) - (nb - na);
                        rejected += (unsigned)removed;
                    }
                    const int a_old = a, b_old = b;
                    if (active) {
                        a = na;
                        b = nb;
                        if (removed == 0) active = false;  // combine.rs:80-82
                    }
                    // ---- moments of the survivors
                    if (it > 0 && __any_sync(FULL, removed > 0)) {
                        double r1 = 0.0, r2 = 0.0;
                        if (removed > 0) {
#pragma unroll 1
                            for (int g = a_old + l; g < a; g += DL) {
                                const double d = (double)S(g) - c0;
                                r1 += d;
                                r2 = fma(d, d, r2);
                            }
#pragma unroll 1
                            for (int g = b + l; g < b_old; g += DL) {
                                const double d = (double)S(g) - c0;
                                r1 += d;
                                r2 = fma(d, d, r2);
                            }
                        }
                        r1 = quad_sum(r1);
                        r2 = quad_sum(r2);
                        D1 -= r1;
                        D2 -= r2;
                    }
                }
                const bool first = it == 0;
                // a downdate that left less than 1/1024 of the last fresh D2 has cancelled ~10 bits
                const bool stale = !first && active && D2 * 1024.0 < D2_fresh;
                if (first || __any_sync(FULL, stale)) {
                    const int a_loc = a - R * l, b_loc = b - R * l;
                    double m1[2] = {0.0, 0.0}, m2[2] = {0.0, 0.0};
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const bool in = a_loc <= j && j < b_loc;
                        const double d = (double)(in ? s[j] : med) - c0;
                        m1[j & 1] += d;
                        m2[j & 1] = fma(d, d, m2[j & 1]);
                    }
                    const double f1 = quad_sum(m1[0] + m1[1]);
                    const double f2 = quad_sum(m2[0] + m2[1]);
                    if (first || stale) {
                        D1 = f1;
                        D2 = f2;
                        D2_fresh = f2;
                    }
                }
                if (it + 1 >= max_iter || !__any_sync(FULL, active)) break;
                // ---- center / sigma of the next iteration (combine.rs:49-61)
                {
                    const int len = max(b - a, 1);
                    const double mean = abk::div_small((double)len * c0 + D1, len);
                    const double dm = mean - c0;
                    double ss = (D2 - 2.0 * dm * D1) + (double)len * dm * dm;
                    ss = fmax(ss, 0.0);
                    if (active) {
                        center = (float)mean;
                        // (float)max(sd, 1e-10) == max((float)sd, (float)1e-10): rounding is monotone
                        sigma = fmaxf((float)sqrt(abk::div_small(ss, len > 2 ? len - 1 : 1)), 1e-10f);
                    }
                }
            }
            const int len = b - a;
            if (n >= 2) {
                if (len == 0) result = isfinite(last_center) ? last_center : 0.0f;  // combine.rs:85-88
                else result = (float)abk::div_small((double)len * c0 + D1, len);
            }
        }
        if (in_image && l == 0) {
            __stcs(out + pix, result);
            local_rej += rejected;
        }
    }
    cp_async_wait_all();
    const unsigned w = __reduce_add_sync(FULL, local_rej);
    if (lane == 0 && w) atomicAdd(rejected_total, (unsigned long long)w);
}

}  // namespace

template <int R, bool PRE>
static ab_status launch_deep_r(ab_ctx *ctx, const FrameTable &ft, int n_frames, size_t npix, float sl,
                               float sh, unsigned max_iter, const DeepPre &pre, float *d_out,
                               unsigned long long *d_rej, cudaStream_t stream) {
    using G = Geom<R>;
    auto kern = stack_deep_kernel<R, PRE>;
    AB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
    int per_sm = 0;
    AB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DTHREADS, G::SMEM));
    if (per_sm < 1) per_sm = 1;
    const size_t n_tiles = (npix + TILE_PX - 1) / TILE_PX;
    size_t grid = (size_t)ctx->sm_count * per_sm;  // persistent: one wave, tiles strided
    if (grid > n_tiles) grid = n_tiles;
    if (grid < 1) grid = 1;
    int aligned16 = 1;
    for (int f = 0; f < n_frames; ++f)
        if (reinterpret_cast<uintptr_t>(ft.p[f]) & 15u) aligned16 = 0;
    if (pre.peer_frames) aligned16 |= 2;
    kern<<<(unsigned)grid, DTHREADS, G::SMEM, stream>>>(ft, n_frames, npix, sl, sh, max_iter,
                                                       CeConst{1, -1}, pre, aligned16, d_out, d_rej);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

ab_status ab_stack_launch_deep(ab_ctx *ctx, const FrameTable &ft, int n_frames, size_t npix,
                               float sl, float sh, unsigned max_iter, const DeepPre &pre,
                               float *d_out, unsigned long long *d_rej, cudaStream_t stream) {
    const bool use_pre = pre.bias || pre.dark || pre.median_only;
    if (n_frames < 1 || n_frames > AB_MAX_FRAMES)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "deep stack: %d frames", n_frames);
#define AB_DEEP_CASE(r)                                                                              \
    return use_pre ? launch_deep_r<r, true>(ctx, ft, n_frames, npix, sl, sh, max_iter, pre, d_out,   \
                                            d_rej, stream)                                           \
                   : launch_deep_r<r, false>(ctx, ft, n_frames, npix, sl, sh, max_iter, pre, d_out,  \
                                             d_rej, stream);
    if (n_frames <= 32) { AB_DEEP_CASE(8) }
    if (n_frames <= 64) { AB_DEEP_CASE(16) }
    if (n_frames <= 128) { AB_DEEP_CASE(32) }
    AB_DEEP_CASE(64)
#undef AB_DEEP_CASE
}
