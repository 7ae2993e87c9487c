// preview.cu -- preview/export tail, LRGB merge, star mask and masked stretch for sm_100a
// (the step after the hot path in every command, and the Stretch step's
// default "masked" mode).
//
// Reference semantics (paths relative to src-tauri/src/):
//   infra/render/rgb.rs:7-47,49-98        render_rgb (u8), render_rgb_16bit (u16 big-endian)
//   cmd/helpers.rs:204-262                render_rgb_preview (nearest decimation to max_dim)
//   cmd/helpers.rs:264-322                render_rgb_preview_with_stf
//   core/imaging/stf.rs:122-145           make_stf_u8_fn
//   core/compose/lrgb.rs:4-62             apply_lrgb, synthesize_luminance
//   core/imaging/star_mask.rs:39-138      generate_star_mask(_from_detection)
//   core/imaging/masked_stretch.rs:45-258 masked_stretch(_with_mask), compute_luminance,
//                                         normalize_to_01, compute_masked_median, apply_mtf
//
// GPU mapping.  Every step is an HBM sweep.  The RGB8 interleave writes 12 bytes (4 pixels) per
// thread as three aligned 32-bit stores over the linear output index, with float4 plane loads when
// nothing is decimated.  The star mask is one block per star scattering `max` through atomicMax
// on the (non-negative) float bit patterns -- max is order-independent, so the result is the
// reference's.  The masked median is an MSB-first radix select over (pixel, mask) pairs; the
// MTF stretch and the mask blend are one fused in-place sweep per iteration.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include <math_constants.h>

#include "ab_internal.h"

namespace {

constexpr int PV_THREADS = 256;

__device__ __forceinline__ bool finite_f(float v) { return fabsf(v) < CUDART_INF_F; }
__device__ __forceinline__ bool valid_px(float v) { return finite_f(v) && v > 1e-7f; }
__device__ __forceinline__ float clamp01(float x) {  // f32::clamp(0, 1): NaN stays NaN
    x = x < 0.0f ? 0.0f : x;
    return x > 1.0f ? 1.0f : x;
}
__device__ __forceinline__ double clampd01(double x) {
    x = x < 0.0 ? 0.0 : x;
    return x > 1.0 ? 1.0 : x;
}
// Rust `f32 as u8` / `as u16` of a value already in [0, hi] or NaN: truncating, NaN -> 0
__device__ __forceinline__ unsigned trunc_sat(float x, float hi) {
    if (!(x == x) || x <= 0.0f) return 0u;
    return x >= hi ? (unsigned)hi : (unsigned)x;
}
// Rust `f64 as usize` for the preview source index (non-negative by construction)
__device__ __forceinline__ size_t idx_of(double x) { return x > 0.0 ? (size_t)x : (size_t)0; }

struct StfTx {  // stf.rs:122-134
    double inv_range, dmin, shadow, inv_clip, midtone;
};
__device__ __forceinline__ unsigned stf_u8(const StfTx &t, float v) {  // stf.rs:136-144
    if (!valid_px(v)) return 0u;
    const double norm = ((double)v - t.dmin) * t.inv_range;
    const double x = clampd01((norm - t.shadow) * t.inv_clip);
    double y;
    if (x <= 0.0) y = 0.0;
    else if (x >= 1.0) y = 1.0;
    else y = (t.midtone - 1.0) * x / ((2.0 * t.midtone - 1.0) * x - t.midtone);
    const double q = round(y * 255.0);  // half away from zero, then saturating `as u8`
    if (!(q == q) || q <= 0.0) return 0u;
    return q >= 255.0 ? 255u : (unsigned)q;
}
struct StfTx3 {
    StfTx t[3];
};
StfTx make_tx(const ab_stf_params &p, const ab_image_stats &st) {
    StfTx t;
    t.inv_range = 1.0 / std::fmax(st.max - st.min, 1e-30);
    t.dmin = st.min;
    t.shadow = p.shadow;
    t.inv_clip = 1.0 / std::fmax(p.highlight - p.shadow, 1e-15);
    t.midtone = p.midtone;
    return t;
}

unsigned blocks_for(ab_ctx *ctx, size_t items) {
    size_t blocks = (items + PV_THREADS - 1) / PV_THREADS;
    const size_t cap = (size_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks ? blocks : 1);
}

// ---- RGB8 interleave.  STF = false: (v.clamp(0,1) * 255) as u8; true: per-channel STF closure.
// FITS = true: source index == output index (no decimation).
template <bool STF, bool FITS>
__global__ void __launch_bounds__(PV_THREADS)
render_u8_kernel(const float *__restrict__ r, const float *__restrict__ g,
                 const float *__restrict__ b, size_t rows, size_t cols, size_t pw, size_t ph,
                 double y_ratio, double x_ratio, StfTx3 tx, bool vec_ok,
                 unsigned char *__restrict__ out) {
    const size_t n_out = pw * ph;
    const size_t n_groups = (n_out + 3) / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const bool out_al = (reinterpret_cast<uintptr_t>(out) & 3) == 0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_groups; k += stride) {
        const size_t o0 = k * 4;
        const int cnt = (int)min((size_t)4, n_out - o0);
        float rv[4] = {0.f, 0.f, 0.f, 0.f}, gv[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (FITS && vec_ok && cnt == 4) {
            const float4 a = __ldcs(reinterpret_cast<const float4 *>(r) + k);
            const float4 c = __ldcs(reinterpret_cast<const float4 *>(g) + k);
            const float4 d = __ldcs(reinterpret_cast<const float4 *>(b) + k);
            rv[0] = a.x; rv[1] = a.y; rv[2] = a.z; rv[3] = a.w;
            gv[0] = c.x; gv[1] = c.y; gv[2] = c.z; gv[3] = c.w;
            bv[0] = d.x; bv[1] = d.y; bv[2] = d.z; bv[3] = d.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (q >= cnt) break;
                size_t si = o0 + q;
                if (!FITS) {  // helpers.rs:241-245 / :300-304
                    const size_t dy = (o0 + q) / pw, dx = (o0 + q) - dy * pw;
                    const size_t sy = idx_of(fmin((double)dy * y_ratio, (double)(rows - 1)));
                    const size_t sx = idx_of(fmin((double)dx * x_ratio, (double)(cols - 1)));
                    si = sy * cols + sx;
                }
                rv[q] = r[si]; gv[q] = g[si]; bv[q] = b[si];
            }
        }
        unsigned char px[12];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (STF) {
                px[3 * q] = (unsigned char)stf_u8(tx.t[0], rv[q]);
                px[3 * q + 1] = (unsigned char)stf_u8(tx.t[1], gv[q]);
                px[3 * q + 2] = (unsigned char)stf_u8(tx.t[2], bv[q]);
            } else {
                px[3 * q] = (unsigned char)trunc_sat(clamp01(rv[q]) * 255.0f, 255.0f);
                px[3 * q + 1] = (unsigned char)trunc_sat(clamp01(gv[q]) * 255.0f, 255.0f);
                px[3 * q + 2] = (unsigned char)trunc_sat(clamp01(bv[q]) * 255.0f, 255.0f);
            }
        }
        unsigned char *dst = out + o0 * 3;
        if (out_al && cnt == 4) {
            unsigned w[3];
#pragma unroll
            for (int j = 0; j < 3; ++j)
                w[j] = (unsigned)px[4 * j] | ((unsigned)px[4 * j + 1] << 8) |
                       ((unsigned)px[4 * j + 2] << 16) | ((unsigned)px[4 * j + 3] << 24);
            unsigned *d32 = reinterpret_cast<unsigned *>(dst);
            d32[0] = w[0]; d32[1] = w[1]; d32[2] = w[2];
        } else {
            for (int j = 0; j < 3 * cnt; ++j) dst[j] = px[j];
        }
    }
}

// ---- RGB16 big-endian interleave (rgb.rs:49-98): 2 pixels = 12 bytes per thread
__global__ void __launch_bounds__(PV_THREADS)
render_u16be_kernel(const float *__restrict__ r, const float *__restrict__ g,
                    const float *__restrict__ b, size_t len, unsigned char *__restrict__ out) {
    const size_t n_groups = (len + 1) / 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const bool out_al = (reinterpret_cast<uintptr_t>(out) & 3) == 0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_groups; k += stride) {
        const size_t i0 = k * 2;
        const int cnt = (int)min((size_t)2, len - i0);
        unsigned char px[12];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const size_t i = i0 + (q < cnt ? q : 0);
            const float ch[3] = {r[i], g[i], b[i]};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const unsigned v = trunc_sat(clamp01(ch[c]) * 65535.0f, 65535.0f);
                px[6 * q + 2 * c] = (unsigned char)(v >> 8);
                px[6 * q + 2 * c + 1] = (unsigned char)(v & 255u);
            }
        }
        unsigned char *dst = out + i0 * 6;
        if (out_al && cnt == 2) {
            unsigned *d32 = reinterpret_cast<unsigned *>(dst);
#pragma unroll
            for (int j = 0; j < 3; ++j)
                d32[j] = (unsigned)px[4 * j] | ((unsigned)px[4 * j + 1] << 8) |
                         ((unsigned)px[4 * j + 2] << 16) | ((unsigned)px[4 * j + 3] << 24);
        } else {
            for (int j = 0; j < 6 * cnt; ++j) dst[j] = px[j];
        }
    }
}

// ---- apply_lrgb, lrgb.rs:26-43 (in place on r, g, b)
__device__ __forceinline__ void lrgb_px(float lum_new, float &rv, float &gv, float &bv, float lw,
                                        float cw) {
    const float lum_old = rv * 0.2126f + gv * 0.7152f + bv * 0.0722f;
    if (lum_old < 1e-10f) {
        const float blended = lum_new * lw;
        rv = gv = bv = blended;
        return;
    }
    const float ratio = (lum_new * lw + lum_old * (1.0f - lw)) / lum_old;
    rv = clamp01(rv * ratio * cw + lum_new * (1.0f - cw));
    gv = clamp01(gv * ratio * cw + lum_new * (1.0f - cw));
    bv = clamp01(bv * ratio * cw + lum_new * (1.0f - cw));
}
__global__ void __launch_bounds__(PV_THREADS)
lrgb_kernel(const float *__restrict__ l, float *__restrict__ r, float *__restrict__ g,
            float *__restrict__ b, size_t len, float lw, float cw, bool vec_ok) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n4 = vec_ok ? len / 4 : 0;
    for (size_t i = tid; i < n4; i += stride) {
        const float4 lv = __ldcs(reinterpret_cast<const float4 *>(l) + i);
        float4 a = reinterpret_cast<float4 *>(r)[i], c = reinterpret_cast<float4 *>(g)[i],
               d = reinterpret_cast<float4 *>(b)[i];
        lrgb_px(lv.x, a.x, c.x, d.x, lw, cw);
        lrgb_px(lv.y, a.y, c.y, d.y, lw, cw);
        lrgb_px(lv.z, a.z, c.z, d.z, lw, cw);
        lrgb_px(lv.w, a.w, c.w, d.w, lw, cw);
        reinterpret_cast<float4 *>(r)[i] = a;
        reinterpret_cast<float4 *>(g)[i] = c;
        reinterpret_cast<float4 *>(b)[i] = d;
    }
    for (size_t i = n4 * 4 + tid; i < len; i += stride) lrgb_px(l[i], r[i], g[i], b[i], lw, cw);
}

// ---- synthesize_luminance (lrgb.rs:57-59) / compute_luminance (masked_stretch.rs:152-157)
template <bool FINITE_ONLY>
__device__ __forceinline__ float lum_px(float rv, float gv, float bv) {
    if (FINITE_ONLY) {
        const float rn = finite_f(rv) ? rv : 0.0f, gn = finite_f(gv) ? gv : 0.0f,
                    bn = finite_f(bv) ? bv : 0.0f;
        return 0.2126f * rn + 0.7152f * gn + 0.0722f * bn;
    }
    return rv * 0.2126f + gv * 0.7152f + bv * 0.0722f;
}
template <bool FINITE_ONLY>
__global__ void __launch_bounds__(PV_THREADS)
luminance_kernel(const float *__restrict__ r, const float *__restrict__ g,
                 const float *__restrict__ b, size_t len, float *__restrict__ out, bool vec_ok) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n4 = vec_ok ? len / 4 : 0;
    for (size_t i = tid; i < n4; i += stride) {
        const float4 a = __ldcs(reinterpret_cast<const float4 *>(r) + i);
        const float4 c = __ldcs(reinterpret_cast<const float4 *>(g) + i);
        const float4 d = __ldcs(reinterpret_cast<const float4 *>(b) + i);
        float4 o;
        o.x = lum_px<FINITE_ONLY>(a.x, c.x, d.x);
        o.y = lum_px<FINITE_ONLY>(a.y, c.y, d.y);
        o.z = lum_px<FINITE_ONLY>(a.z, c.z, d.z);
        o.w = lum_px<FINITE_ONLY>(a.w, c.w, d.w);
        __stcs(reinterpret_cast<float4 *>(out) + i, o);
    }
    for (size_t i = n4 * 4 + tid; i < len; i += stride) out[i] = lum_px<FINITE_ONLY>(r[i], g[i], b[i]);
}

// ---- star mask (star_mask.rs:62-108): per-star parameters computed on the host in f64
struct MaskStar {
    double x, y, r2_inner, r2_outer, fade_range;
    unsigned y_min, y_max, x_min, x_max;  // inclusive; y_min > y_max: empty
};
__global__ void __launch_bounds__(PV_THREADS)
star_mask_kernel(const MaskStar *__restrict__ stars, size_t w, float *__restrict__ mask) {
    const MaskStar s = stars[blockIdx.x];
    if (s.y_min > s.y_max || s.x_min > s.x_max) return;
    const unsigned bw = s.x_max - s.x_min + 1, bh = s.y_max - s.y_min + 1;
    const unsigned long long n = (unsigned long long)bw * bh;
    for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned py = s.y_min + (unsigned)(i / bw), px = s.x_min + (unsigned)(i % bw);
        const double dx = (double)px - s.x, dy = (double)py - s.y;
        const double d2 = dx * dx + dy * dy;
        float val;
        if (d2 <= s.r2_inner) {
            val = 1.0f;
        } else if (d2 <= s.r2_outer) {
            const float t = (float)((d2 - s.r2_inner) / s.fade_range);
            const float smooth = t * t * (3.0f - 2.0f * t);
            val = 1.0f - smooth;
        } else {
            continue;
        }
        // `if val > current { current = val }` over all stars == max; val >= 0 so the int order
        // of the bit patterns is the float order
        if (val > 0.0f) atomicMax(reinterpret_cast<int *>(mask) + ((size_t)py * w + px), __float_as_int(val));
    }
}
// luminance protection (star_mask.rs:110-128) + coverage count (:131)
__global__ void __launch_bounds__(PV_THREADS)
mask_finish_kernel(const float *__restrict__ img, float *__restrict__ mask, size_t len, bool protect,
                   float ceiling, float inv_range, unsigned long long *__restrict__ covered) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned cnt = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        float m = mask[i];
        if (protect) {
            const float pixel = img[i];
            if (pixel > ceiling && m < 1.0f) {
                const float excess = clamp01((pixel - ceiling) * inv_range);
                const float smooth = excess * excess * (3.0f - 2.0f * excess);
                if (smooth > m) {
                    m = smooth;
                    mask[i] = m;
                }
            }
        }
        cnt += m > 0.01f;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(covered, (unsigned long long)cnt);
}

// ---- masked stretch pieces
// normalize_to_01, masked_stretch.rs:192-209
__global__ void __launch_bounds__(PV_THREADS)
norm01_kernel(const float *__restrict__ img, float *__restrict__ out, size_t len, float dmin,
              float inv, bool zero_all) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const float v = img[i];
        out[i] = (zero_all || !finite_f(v) || v <= 0.0f) ? 0.0f : clamp01((v - dmin) * inv);
    }
}
// one radix pass of compute_masked_median (masked_stretch.rs:211-228): keys are the bit patterns
// of the qualifying pixels (mask < 0.5, finite, > 0: positive floats, bits monotone)
__global__ void __launch_bounds__(PV_THREADS)
masked_radix_kernel(const float *__restrict__ img, const float *__restrict__ mask, size_t len,
                    unsigned prefix, unsigned prefix_mask, int shift,
                    unsigned *__restrict__ hist256) {
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = 0;  // PV_THREADS == 256
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const float v = img[i];
        if (!(mask[i] < 0.5f && finite_f(v) && v > 0.0f)) continue;
        const unsigned key = __float_as_uint(v);
        if ((key & prefix_mask) == prefix) atomicAdd(&sh[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist256[threadIdx.x], sh[threadIdx.x]);
}
// apply_mtf (masked_stretch.rs:238-254) fused with the mask blend (:96-102), in place
__global__ void __launch_bounds__(PV_THREADS)
mtf_blend_kernel(float *__restrict__ work, const float *__restrict__ mask, size_t len, float m,
                 float protection) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const float x = work[i];
        float stretched;
        if (x <= 0.0f) {
            stretched = 0.0f;
        } else if (x >= 1.0f) {
            stretched = 1.0f;
        } else {
            const float d = (2.0f * m - 1.0f) * x - m;
            stretched = fabsf(d) < 1e-10f ? x : clamp01((m - 1.0f) * x / d);
        }
        const float blend = mask[i] * protection;
        work[i] = x * blend + stretched * (1.0f - blend);
    }
}
__global__ void __launch_bounds__(PV_THREADS)
clamp01_kernel(float *__restrict__ data, size_t len) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride)
        data[i] = clamp01(data[i]);
}

// ------------------------------------------------------------------------------ host logic
void preview_size(size_t rows, size_t cols, size_t max_dim, size_t *pw, size_t *ph, bool *fits) {
    *fits = max_dim == 0 || (rows <= max_dim && cols <= max_dim);
    if (*fits) {
        *pw = cols;
        *ph = rows;
        return;
    }
    const double scale = (double)max_dim / (double)std::max(rows, cols);  // helpers.rs:226-228
    *pw = (size_t)std::fmax(std::round((double)cols * scale), 1.0);
    *ph = (size_t)std::fmax(std::round((double)rows * scale), 1.0);
}

ab_status render_u8_on(ab_ctx *ctx, const float *r, const float *g, const float *b, size_t rows,
                       size_t cols, size_t max_dim, const ab_stf_params *stf3,
                       const ab_image_stats *st3, unsigned char *out, cudaStream_t s) {
    if (rows == 0 || cols == 0) return AB_OK;
    size_t pw, ph;
    bool fits;
    preview_size(rows, cols, max_dim, &pw, &ph, &fits);
    const double y_ratio = fits ? 1.0 : (double)rows / (double)ph;
    const double x_ratio = fits ? 1.0 : (double)cols / (double)pw;
    StfTx3 tx{};
    if (stf3)
        for (int c = 0; c < 3; ++c) tx.t[c] = make_tx(stf3[c], st3[c]);
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(g) |
                          reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    const unsigned grid = blocks_for(ctx, (pw * ph + 3) / 4);
#define AB_RENDER(STF, FITS)                                                                  \
    render_u8_kernel<STF, FITS><<<grid, PV_THREADS, 0, s>>>(r, g, b, rows, cols, pw, ph, y_ratio, \
                                                            x_ratio, tx, vec_ok, out)
    if (stf3) {
        if (fits) AB_RENDER(true, true);
        else AB_RENDER(true, false);
    } else {
        if (fits) AB_RENDER(false, true);
        else AB_RENDER(false, false);
    }
#undef AB_RENDER
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

ab_status star_mask_on(ab_ctx *ctx, const float *d_img, size_t h, size_t w, const ab_star *stars,
                       size_t n_stars, const ab_star_mask_config *cfg, float *d_mask,
                       ab_star_mask_info *info, cudaStream_t s) {
    const size_t len = h * w;
    info->stars_masked = 0;
    info->coverage_fraction = 0.0;
    if (len == 0) return AB_OK;
    if (h >= ((size_t)1 << 32) || w >= ((size_t)1 << 32))
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "image too large for the star mask");
    std::vector<MaskStar> ms;
    auto usz = [](double x) -> size_t {  // Rust `f64 as usize`
        if (!(x == x) || x <= 0.0) return 0;
        if (x >= 18446744073709551615.0) return (size_t)-1;
        return (size_t)x;
    };
    for (size_t i = 0; i < n_stars; ++i) {
        const ab_star &st = stars[i];
        if (!(st.fwhm >= cfg->min_fwhm && st.fwhm <= cfg->max_fwhm)) continue;  // star_mask.rs:56
        ++info->stars_masked;
        const double radius = st.fwhm * cfg->growth_factor;
        const double soft_radius = radius + cfg->softness;
        MaskStar m;
        m.x = st.x;
        m.y = st.y;
        const size_t y_min = usz(std::fmax(std::floor(st.y - soft_radius), 0.0));
        const size_t y_max = std::min(usz(std::ceil(st.y + soft_radius)), h - 1);
        const size_t x_min = usz(std::fmax(std::floor(st.x - soft_radius), 0.0));
        const size_t x_max = std::min(usz(std::ceil(st.x + soft_radius)), w - 1);
        if (y_min > y_max || x_min > x_max) continue;  // empty ranges (star off the frame)
        m.y_min = (unsigned)y_min; m.y_max = (unsigned)y_max;
        m.x_min = (unsigned)x_min; m.x_max = (unsigned)x_max;
        m.r2_inner = radius * radius;
        m.r2_outer = soft_radius * soft_radius;
        m.fade_range = std::fmax(m.r2_outer - m.r2_inner, 1e-10);
        ms.push_back(m);
    }
    AB_CUDA(ctx, cudaMemsetAsync(d_mask, 0, len * sizeof(float), s));
    unsigned long long *d_cov = reinterpret_cast<unsigned long long *>(ctx->d_small);
    AB_CUDA(ctx, cudaMemsetAsync(d_cov, 0, sizeof(unsigned long long), s));
    if (!ms.empty()) {
        ab_status st = ab_scratch_reserve(ctx, 2, ms.size() * sizeof(MaskStar));
        if (st != AB_OK) return st;
        AB_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch[2], ms.data(), ms.size() * sizeof(MaskStar),
                                     cudaMemcpyHostToDevice, s));
        // `ms` must outlive the pageable copy: cudaMemcpyAsync from pageable memory returns
        // after the source has been staged
        for (size_t off = 0; off < ms.size(); off += 65535) {
            const unsigned nb = (unsigned)std::min<size_t>(65535, ms.size() - off);
            star_mask_kernel<<<nb, PV_THREADS, 0, s>>>(
                static_cast<const MaskStar *>(ctx->d_scratch[2]) + off, w, d_mask);
            AB_LAUNCH_CHECK(ctx);
        }
    }
    const float ceiling = (float)cfg->luminance_ceiling;
    const float inv_range = ceiling < 1.0f ? 1.0f / (1.0f - ceiling) : 1.0f;
    mask_finish_kernel<<<blocks_for(ctx, len), PV_THREADS, 0, s>>>(
        d_img, d_mask, len, cfg->luminance_protect != 0, ceiling, inv_range, d_cov);
    AB_LAUNCH_CHECK(ctx);
    unsigned long long *h_cov = reinterpret_cast<unsigned long long *>(ctx->h_small);
    AB_CUDA(ctx, cudaMemcpyAsync(h_cov, d_cov, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    AB_CUDA(ctx, cudaStreamSynchronize(s));
    info->coverage_fraction = (double)*h_cov / (double)len;
    return AB_OK;
}

// compute_masked_median: upper median (element len/2) of the qualifying pixels, 0.0 if none
ab_status masked_median_on(ab_ctx *ctx, const float *d_img, const float *d_mask, size_t len,
                           cudaStream_t s, double *out) {
    *out = 0.0;
    if (len == 0) return AB_OK;
    unsigned *d_hist = reinterpret_cast<unsigned *>(static_cast<char *>(ctx->d_small) + 65536);
    unsigned *h_hist = reinterpret_cast<unsigned *>(static_cast<char *>(ctx->h_small) + 65536);
    unsigned prefix = 0, pmask = 0;
    unsigned long long k = 0;
    const unsigned grid = blocks_for(ctx, len / 8 + 1);
    for (int shift = 24; shift >= 0; shift -= 8) {
        AB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, 256 * sizeof(unsigned), s));
        masked_radix_kernel<<<grid, PV_THREADS, 0, s>>>(d_img, d_mask, len, prefix, pmask, shift, d_hist);
        AB_LAUNCH_CHECK(ctx);
        AB_CUDA(ctx, cudaMemcpyAsync(h_hist, d_hist, 256 * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
        AB_CUDA(ctx, cudaStreamSynchronize(s));
        if (shift == 24) {
            unsigned long long n = 0;
            for (int b = 0; b < 256; ++b) n += h_hist[b];
            if (n == 0) return AB_OK;
            k = n / 2;
        }
        unsigned long long cum = 0;
        int b = 0;
        for (; b < 256; ++b) {
            if (cum + h_hist[b] > k) break;
            cum += h_hist[b];
        }
        if (b == 256) return ab_fail(ctx, AB_ERR_CUDA, "masked median: rank out of range");
        k -= cum;
        prefix |= (unsigned)b << shift;
        pmask |= 0xffu << shift;
    }
    float r;
    std::memcpy(&r, &prefix, 4);
    *out = (double)r;
    return AB_OK;
}

double mtf_balance(double median, double target) {  // masked_stretch.rs:230-236
    const double denom = 2.0 * target * median - target - median;
    if (std::fabs(denom) < 1e-15) return 0.5;
    const double v = median * (target - 1.0) / denom;
    return v < 0.0001 ? 0.0001 : (v > 0.9999 ? 0.9999 : v);
}

// masked_stretch_with_mask, masked_stretch.rs:59-117; d_out is the working image
ab_status masked_stretch_on(ab_ctx *ctx, const float *d_img, size_t len, const float *d_mask,
                            const ab_masked_stretch_config *cfg, float *d_out,
                            ab_masked_stretch_info *info, cudaStream_t s) {
    info->iterations_run = 0;
    info->final_background = 0.0;
    info->converged = 0;
    if (len == 0) return AB_OK;
    double mn, mx, sum;
    unsigned long long n;
    ab_status st = ab_scan_on(ctx, d_img, len, s, &mn, &mx, &sum, &n);  // stats.min / stats.max
    if (st != AB_OK) return st;
    if (n == 0) mn = mx = 0.0;  // ImageStats::default()
    const float range = (float)(mx - mn);
    const bool zero_all = range < 1e-10f;
    const unsigned grid = blocks_for(ctx, len / 4 + 1);
    norm01_kernel<<<grid, PV_THREADS, 0, s>>>(d_img, d_out, len, (float)mn,
                                              zero_all ? 0.0f : 1.0f / range, zero_all);
    AB_LAUNCH_CHECK(ctx);
    const float protection = (float)cfg->protection_amount;
    double prev_bg;
    if ((st = masked_median_on(ctx, d_out, d_mask, len, s, &prev_bg)) != AB_OK) return st;
    for (uint64_t it = 0; it < cfg->iterations; ++it) {
        info->iterations_run = it + 1;
        double bg = prev_bg;
        // the working image only changes after a stretch: iteration 0 re-measures what prev_bg holds
        if (it > 0 && (st = masked_median_on(ctx, d_out, d_mask, len, s, &bg)) != AB_OK) return st;
        const bool at_target = std::fabs(bg - cfg->target_background) < cfg->convergence_threshold;
        const bool stagnated = it > 0 && std::fabs(bg - prev_bg) < cfg->convergence_threshold * 0.1;
        if (at_target) {
            info->converged = 1;
            break;
        }
        if (stagnated) break;
        const float m = (float)mtf_balance(bg, cfg->target_background);
        mtf_blend_kernel<<<grid, PV_THREADS, 0, s>>>(d_out, d_mask, len, m, protection);
        AB_LAUNCH_CHECK(ctx);
        prev_bg = bg;
    }
    if ((st = masked_median_on(ctx, d_out, d_mask, len, s, &info->final_background)) != AB_OK) return st;
    clamp01_kernel<<<grid, PV_THREADS, 0, s>>>(d_out, len);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

ab_star_mask_config mask_cfg_of(const ab_masked_stretch_config *c) {  // masked_stretch.rs:49-55
    ab_star_mask_config m;
    m.growth_factor = c->mask_growth;
    m.softness = c->mask_softness;
    m.detection_sigma = 5.0;
    m.min_fwhm = 1.5;
    m.max_fwhm = 30.0;
    m.luminance_protect = c->luminance_protect;
    m.luminance_ceiling = c->luminance_ceiling;
    return m;
}

ab_status generate_mask_on(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols,
                           const ab_star_mask_config *cfg, float *d_mask, ab_star_mask_info *info,
                           cudaStream_t s) {
    std::vector<ab_star> stars;
    double bm, bs;
    ab_status st = ab_detect_stars_on(ctx, d_img, rows, cols, cfg->detection_sigma, s, stars, &bm, &bs);
    if (st != AB_OK) return st;
    return star_mask_on(ctx, d_img, rows, cols, stars.data(), stars.size(), cfg, d_mask, info, s);
}

// scratch slot 0 carved into 256-byte aligned planes
struct Planes {
    char *base = nullptr;
    size_t plane = 0;
    float *at(int i) const { return reinterpret_cast<float *>(base + (size_t)i * plane); }
};
ab_status planes_reserve(ab_ctx *ctx, size_t len, int n, Planes *p) {
    p->plane = (len * sizeof(float) + 255) / 256 * 256;
    ab_status st = ab_scratch_reserve(ctx, 0, p->plane * n + 256);
    if (st != AB_OK) return st;
    p->base = static_cast<char *>(ctx->d_scratch[0]);
    return AB_OK;
}

}  // namespace

extern "C" {

ab_status ab_preview_size(size_t rows, size_t cols, size_t max_dim, size_t *pw, size_t *ph) {
    if (!pw || !ph) return AB_ERR_INVALID;
    bool fits;
    preview_size(rows, cols, max_dim, pw, ph, &fits);
    return AB_OK;
}

// ---------------------------------------------------------------- RGB8 / RGB16 interleave
ab_status ab_render_rgb_u8_dev(ab_ctx *ctx, const float *d_r, const float *d_g, const float *d_b,
                               size_t rows, size_t cols, size_t max_dim, uint8_t *d_out,
                               void *stream) {
    if (!ctx || (rows * cols && (!d_r || !d_g || !d_b || !d_out))) return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    return render_u8_on(ctx, d_r, d_g, d_b, rows, cols, max_dim, nullptr, nullptr, d_out,
                        (cudaStream_t)stream);
}
ab_status ab_render_rgb_stf_u8_dev(ab_ctx *ctx, const float *d_r, const float *d_g,
                                   const float *d_b, size_t rows, size_t cols,
                                   const ab_stf_params *stf3, const ab_image_stats *stats3,
                                   size_t max_dim, uint8_t *d_out, void *stream) {
    if (!ctx || !stf3 || !stats3 || (rows * cols && (!d_r || !d_g || !d_b || !d_out)))
        return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    return render_u8_on(ctx, d_r, d_g, d_b, rows, cols, max_dim, stf3, stats3, d_out,
                        (cudaStream_t)stream);
}
static ab_status render_host(ab_ctx *ctx, const float *r, const float *g, const float *b,
                             size_t rows, size_t cols, size_t max_dim, const ab_stf_params *stf3,
                             const ab_image_stats *st3, int mode16, uint8_t *out) {
    const size_t len = rows * cols;
    if (len == 0) return AB_OK;
    CtxLock lk(ctx->mu);
    DeviceGuard gd(ctx->device);
    size_t pw, ph;
    bool fits;
    preview_size(rows, cols, mode16 ? 0 : max_dim, &pw, &ph, &fits);
    const size_t out_bytes = pw * ph * (mode16 ? 6 : 3);
    Planes p;
    const int out_planes = (int)((out_bytes + len * 4 - 1) / (len * 4));
    ab_status st = planes_reserve(ctx, len, 3 + out_planes, &p);
    if (st != AB_OK) return st;
    const float *in[3] = {r, g, b};
    for (int c = 0; c < 3; ++c)
        AB_CUDA(ctx, cudaMemcpyAsync(p.at(c), in[c], len * 4, cudaMemcpyHostToDevice, ctx->stream));
    unsigned char *d_out = reinterpret_cast<unsigned char *>(p.at(3));
    if (mode16) {
        render_u16be_kernel<<<blocks_for(ctx, (len + 1) / 2), PV_THREADS, 0, ctx->stream>>>(
            p.at(0), p.at(1), p.at(2), len, d_out);
        AB_LAUNCH_CHECK(ctx);
    } else {
        st = render_u8_on(ctx, p.at(0), p.at(1), p.at(2), rows, cols, max_dim, stf3, st3, d_out,
                          ctx->stream);
        if (st != AB_OK) return st;
    }
    AB_CUDA(ctx, cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}
ab_status ab_render_rgb_u8(ab_ctx *ctx, const float *r, const float *g, const float *b, size_t rows,
                           size_t cols, size_t max_dim, uint8_t *out) {
    if (!ctx || (rows * cols && (!r || !g || !b || !out))) return AB_ERR_INVALID;
    return render_host(ctx, r, g, b, rows, cols, max_dim, nullptr, nullptr, 0, out);
}
ab_status ab_render_rgb_stf_u8(ab_ctx *ctx, const float *r, const float *g, const float *b,
                               size_t rows, size_t cols, const ab_stf_params *stf3,
                               const ab_image_stats *stats3, size_t max_dim, uint8_t *out) {
    if (!ctx || !stf3 || !stats3 || (rows * cols && (!r || !g || !b || !out))) return AB_ERR_INVALID;
    return render_host(ctx, r, g, b, rows, cols, max_dim, stf3, stats3, 0, out);
}
ab_status ab_render_rgb_u16be_dev(ab_ctx *ctx, const float *d_r, const float *d_g,
                                  const float *d_b, size_t len, uint8_t *d_out, void *stream) {
    if (!ctx || (len && (!d_r || !d_g || !d_b || !d_out))) return AB_ERR_INVALID;
    if (len == 0) return AB_OK;
    DeviceGuard g(ctx->device);
    render_u16be_kernel<<<blocks_for(ctx, (len + 1) / 2), PV_THREADS, 0, (cudaStream_t)stream>>>(
        d_r, d_g, d_b, len, d_out);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}
ab_status ab_render_rgb_u16be(ab_ctx *ctx, const float *r, const float *g, const float *b,
                              size_t len, uint8_t *out) {
    if (!ctx || (len && (!r || !g || !b || !out))) return AB_ERR_INVALID;
    return render_host(ctx, r, g, b, 1, len, 0, nullptr, nullptr, 1, out);
}

// ---------------------------------------------------------------- LRGB / luminance
ab_status ab_apply_lrgb_dev(ab_ctx *ctx, const float *d_l, float *d_r, float *d_g, float *d_b,
                            size_t len, float lightness_weight, float chrominance_weight,
                            void *stream) {
    if (!ctx || (len && (!d_l || !d_r || !d_g || !d_b))) return AB_ERR_INVALID;
    if (len == 0) return AB_OK;
    DeviceGuard g(ctx->device);
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(d_l) | reinterpret_cast<uintptr_t>(d_r) |
                          reinterpret_cast<uintptr_t>(d_g) | reinterpret_cast<uintptr_t>(d_b)) & 15) == 0;
    lrgb_kernel<<<blocks_for(ctx, len / 4 + 1), PV_THREADS, 0, (cudaStream_t)stream>>>(
        d_l, d_r, d_g, d_b, len, lightness_weight, chrominance_weight, vec_ok);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}
ab_status ab_apply_lrgb(ab_ctx *ctx, const float *l, float *r, float *g, float *b, size_t len,
                        float lightness_weight, float chrominance_weight) {
    if (!ctx || (len && (!l || !r || !g || !b))) return AB_ERR_INVALID;
    if (len == 0) return AB_OK;
    CtxLock lk(ctx->mu);
    DeviceGuard gd(ctx->device);
    Planes p;
    ab_status st = planes_reserve(ctx, len, 4, &p);
    if (st != AB_OK) return st;
    const float *in[4] = {l, r, g, b};
    for (int c = 0; c < 4; ++c)
        AB_CUDA(ctx, cudaMemcpyAsync(p.at(c), in[c], len * 4, cudaMemcpyHostToDevice, ctx->stream));
    lrgb_kernel<<<blocks_for(ctx, len / 4 + 1), PV_THREADS, 0, ctx->stream>>>(
        p.at(0), p.at(1), p.at(2), p.at(3), len, lightness_weight, chrominance_weight, true);
    AB_LAUNCH_CHECK(ctx);
    float *o[3] = {r, g, b};
    for (int c = 0; c < 3; ++c)
        AB_CUDA(ctx, cudaMemcpyAsync(o[c], p.at(c + 1), len * 4, cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}
ab_status ab_luminance_dev(ab_ctx *ctx, const float *d_r, const float *d_g, const float *d_b,
                           size_t len, int finite_only, float *d_out, void *stream) {
    if (!ctx || (len && (!d_r || !d_g || !d_b || !d_out))) return AB_ERR_INVALID;
    if (len == 0) return AB_OK;
    DeviceGuard g(ctx->device);
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(d_r) | reinterpret_cast<uintptr_t>(d_g) |
                          reinterpret_cast<uintptr_t>(d_b) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0;
    const unsigned grid = blocks_for(ctx, len / 4 + 1);
    if (finite_only)
        luminance_kernel<true><<<grid, PV_THREADS, 0, (cudaStream_t)stream>>>(d_r, d_g, d_b, len, d_out, vec_ok);
    else
        luminance_kernel<false><<<grid, PV_THREADS, 0, (cudaStream_t)stream>>>(d_r, d_g, d_b, len, d_out, vec_ok);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}
ab_status ab_luminance(ab_ctx *ctx, const float *r, const float *g, const float *b, size_t len,
                       int finite_only, float *out) {
    if (!ctx || (len && (!r || !g || !b || !out))) return AB_ERR_INVALID;
    if (len == 0) return AB_OK;
    CtxLock lk(ctx->mu);
    DeviceGuard gd(ctx->device);
    Planes p;
    ab_status st = planes_reserve(ctx, len, 4, &p);
    if (st != AB_OK) return st;
    const float *in[3] = {r, g, b};
    for (int c = 0; c < 3; ++c)
        AB_CUDA(ctx, cudaMemcpyAsync(p.at(c), in[c], len * 4, cudaMemcpyHostToDevice, ctx->stream));
    const unsigned grid = blocks_for(ctx, len / 4 + 1);
    if (finite_only)
        luminance_kernel<true><<<grid, PV_THREADS, 0, ctx->stream>>>(p.at(0), p.at(1), p.at(2), len, p.at(3), true);
    else
        luminance_kernel<false><<<grid, PV_THREADS, 0, ctx->stream>>>(p.at(0), p.at(1), p.at(2), len, p.at(3), true);
    AB_LAUNCH_CHECK(ctx);
    AB_CUDA(ctx, cudaMemcpyAsync(out, p.at(3), len * 4, cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}

// ---------------------------------------------------------------- star mask
ab_status ab_star_mask_from_stars_dev(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols,
                                      const ab_star *stars, size_t n_stars,
                                      const ab_star_mask_config *cfg, float *d_mask,
                                      ab_star_mask_info *info, void *stream) {
    if (!ctx || !cfg || !info || (n_stars && !stars) || (rows * cols && (!d_img || !d_mask)))
        return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    return star_mask_on(ctx, d_img, rows, cols, stars, n_stars, cfg, d_mask, info, (cudaStream_t)stream);
}
ab_status ab_generate_star_mask_dev(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols,
                                    const ab_star_mask_config *cfg, float *d_mask,
                                    ab_star_mask_info *info, void *stream) {
    if (!ctx || !cfg || !info || (rows * cols && (!d_img || !d_mask))) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    return generate_mask_on(ctx, d_img, rows, cols, cfg, d_mask, info, (cudaStream_t)stream);
}
static ab_status mask_host(ab_ctx *ctx, const float *img, size_t rows, size_t cols,
                           const ab_star *stars, size_t n_stars, bool detect,
                           const ab_star_mask_config *cfg, float *mask, ab_star_mask_info *info) {
    const size_t len = rows * cols;
    info->stars_masked = 0;
    info->coverage_fraction = 0.0;
    if (len == 0) return AB_OK;
    CtxLock lk(ctx->mu);
    DeviceGuard gd(ctx->device);
    Planes p;
    ab_status st = planes_reserve(ctx, len, 2, &p);
    if (st != AB_OK) return st;
    AB_CUDA(ctx, cudaMemcpyAsync(p.at(0), img, len * 4, cudaMemcpyHostToDevice, ctx->stream));
    st = detect ? generate_mask_on(ctx, p.at(0), rows, cols, cfg, p.at(1), info, ctx->stream)
                : star_mask_on(ctx, p.at(0), rows, cols, stars, n_stars, cfg, p.at(1), info, ctx->stream);
    if (st != AB_OK) return st;
    AB_CUDA(ctx, cudaMemcpyAsync(mask, p.at(1), len * 4, cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}
ab_status ab_star_mask_from_stars(ab_ctx *ctx, const float *img, size_t rows, size_t cols,
                                  const ab_star *stars, size_t n_stars,
                                  const ab_star_mask_config *cfg, float *mask,
                                  ab_star_mask_info *info) {
    if (!ctx || !cfg || !info || (n_stars && !stars) || (rows * cols && (!img || !mask)))
        return AB_ERR_INVALID;
    return mask_host(ctx, img, rows, cols, stars, n_stars, false, cfg, mask, info);
}
ab_status ab_generate_star_mask(ab_ctx *ctx, const float *img, size_t rows, size_t cols,
                                const ab_star_mask_config *cfg, float *mask,
                                ab_star_mask_info *info) {
    if (!ctx || !cfg || !info || (rows * cols && (!img || !mask))) return AB_ERR_INVALID;
    return mask_host(ctx, img, rows, cols, nullptr, 0, true, cfg, mask, info);
}

// ---------------------------------------------------------------- masked stretch
ab_status ab_masked_median_dev(ab_ctx *ctx, const float *d_img, const float *d_mask, size_t len,
                               double *out, void *stream) {
    if (!ctx || !out || (len && (!d_img || !d_mask))) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    return masked_median_on(ctx, d_img, d_mask, len, (cudaStream_t)stream, out);
}
ab_status ab_masked_stretch_with_mask_dev(ab_ctx *ctx, const float *d_img, size_t len,
                                          const float *d_mask, const ab_masked_stretch_config *cfg,
                                          float *d_out, ab_masked_stretch_info *info, void *stream) {
    if (!ctx || !cfg || !info || (len && (!d_img || !d_mask || !d_out))) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    info->stars_masked = 0;
    info->mask_coverage = 0.0;
    return masked_stretch_on(ctx, d_img, len, d_mask, cfg, d_out, info, (cudaStream_t)stream);
}
ab_status ab_masked_stretch_with_mask(ab_ctx *ctx, const float *img, size_t len, const float *mask,
                                      const ab_masked_stretch_config *cfg, float *out,
                                      ab_masked_stretch_info *info) {
    if (!ctx || !cfg || !info || (len && (!img || !mask || !out))) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard gd(ctx->device);
    info->stars_masked = 0;
    info->mask_coverage = 0.0;
    info->iterations_run = 0;
    info->final_background = 0.0;
    info->converged = 0;
    if (len == 0) return AB_OK;
    Planes p;
    ab_status st = planes_reserve(ctx, len, 3, &p);
    if (st != AB_OK) return st;
    AB_CUDA(ctx, cudaMemcpyAsync(p.at(0), img, len * 4, cudaMemcpyHostToDevice, ctx->stream));
    AB_CUDA(ctx, cudaMemcpyAsync(p.at(1), mask, len * 4, cudaMemcpyHostToDevice, ctx->stream));
    st = masked_stretch_on(ctx, p.at(0), len, p.at(1), cfg, p.at(2), info, ctx->stream);
    if (st != AB_OK) return st;
    AB_CUDA(ctx, cudaMemcpyAsync(out, p.at(2), len * 4, cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}
// masked_stretch, masked_stretch.rs:45-57: detect (sigma 5) -> mask -> stretch; d_mask receives
// the mask (may be NULL for the host call: scratch is used)
ab_status ab_masked_stretch_dev(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols,
                                const ab_masked_stretch_config *cfg, float *d_mask, float *d_out,
                                ab_masked_stretch_info *info, void *stream) {
    if (!ctx || !cfg || !info || (rows * cols && (!d_img || !d_mask || !d_out))) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    const ab_star_mask_config mc = mask_cfg_of(cfg);
    ab_star_mask_info mi;
    ab_status st = generate_mask_on(ctx, d_img, rows, cols, &mc, d_mask, &mi, (cudaStream_t)stream);
    if (st != AB_OK) return st;
    st = masked_stretch_on(ctx, d_img, rows * cols, d_mask, cfg, d_out, info, (cudaStream_t)stream);
    info->stars_masked = mi.stars_masked;
    info->mask_coverage = mi.coverage_fraction;
    return st;
}
ab_status ab_masked_stretch(ab_ctx *ctx, const float *img, size_t rows, size_t cols,
                            const ab_masked_stretch_config *cfg, float *out,
                            ab_masked_stretch_info *info) {
    if (!ctx || !cfg || !info || (rows * cols && (!img || !out))) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard gd(ctx->device);
    std::memset(info, 0, sizeof(*info));
    const size_t len = rows * cols;
    if (len == 0) return AB_OK;
    Planes p;
    ab_status st = planes_reserve(ctx, len, 3, &p);
    if (st != AB_OK) return st;
    AB_CUDA(ctx, cudaMemcpyAsync(p.at(0), img, len * 4, cudaMemcpyHostToDevice, ctx->stream));
    const ab_star_mask_config mc = mask_cfg_of(cfg);
    ab_star_mask_info mi;
    st = generate_mask_on(ctx, p.at(0), rows, cols, &mc, p.at(1), &mi, ctx->stream);
    if (st != AB_OK) return st;
    st = masked_stretch_on(ctx, p.at(0), len, p.at(1), cfg, p.at(2), info, ctx->stream);
    if (st != AB_OK) return st;
    info->stars_masked = mi.stars_masked;
    info->mask_coverage = mi.coverage_fraction;
    AB_CUDA(ctx, cudaMemcpyAsync(out, p.at(2), len * 4, cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}

}  // extern "C"
