// pointwise.cu -- one-sweep per-pixel operators for sm_100a (HBM-bound, 128-bit vectorised).
//
// Reference semantics (paths relative to src-tauri/src/):
//   core/imaging/stf.rs:50-120          mtf, StfTransform, apply_stf (u8), apply_stf_f32
//   core/compose/rgb.rs:191-207         apply_stf_inplace (divides by clip_range)
//   core/imaging/stretch.rs:10-46       arcsinh_stretch_with_stats
//   core/compose/channel_blend.rs:13-70 blend_channels
//   core/imaging/scnr.rs:7-53           apply_scnr_inplace
//   core/compose/rgb.rs:127-130, cmd/compose/color.rs:21-49   v * factor
//   core/imaging/curves.rs:17-63        levels
//   core/imaging/curves.rs:65-208       SplineLut (host), apply_curve
//
// Every kernel is a grid-stride sweep: 4 pixels per thread per iteration through one 16-byte
// load per input plane and one 16-byte (or 4-byte for u8) store per output plane, grid = SMs x
// resident blocks.  f64 per-pixel math follows the reference expression by expression
// (-fmad=false); the blend is f32 mul-then-add in weight-list order (no FMA, no tensor cores:
// 3xN with N <= 7 is 1.5 flop/B and a TF32 MMA would break the 1e-5 parity bar).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include <math_constants.h>

#include "ab_internal.h"

namespace {

constexpr int PW_THREADS = 256;
constexpr float PADDING_THRESHOLD = 1e-7f;  // types/constants.rs:6

__device__ __forceinline__ bool valid_px(float v) {
    return (fabsf(v) < CUDART_INF_F) && v > PADDING_THRESHOLD;  // finite && > 1e-7 (NaN fails)
}

__device__ __forceinline__ double clampd(double x, double lo, double hi) {  // f64::clamp
    x = x < lo ? lo : x;
    return x > hi ? hi : x;
}

// stf.rs:50-58
__device__ __forceinline__ double mtf(double x, double m) {
    if (x <= 0.0) return 0.0;
    if (x >= 1.0) return 1.0;
    return (m - 1.0) * x / ((2.0 * m - 1.0) * x - m);
}

struct StfTx {  // stf.rs:60-87 (host-computed scalars)
    double inv_range, dmin, shadow, inv_clip, midtone, clip_range;
};

__device__ __forceinline__ double stf_apply(const StfTx &t, double v) {
    const double norm = (v - t.dmin) * t.inv_range;  // not clamped (G6)
    const double clipped = clampd((norm - t.shadow) * t.inv_clip, 0.0, 1.0);
    return mtf(clipped, t.midtone);
}

// Generic 4-wide sweep: F(float) -> OutT applied to each element.
template <typename OutT, typename F>
__global__ void __launch_bounds__(PW_THREADS)
map1_kernel(const float *__restrict__ in, OutT *__restrict__ out, size_t len, F f) {
    const size_t n4 = len / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(in) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(out) & (4 * sizeof(OutT) - 1)) == 0);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec_ok) {
        for (; i < n4; i += stride) {
            const float4 v = __ldcs(reinterpret_cast<const float4 *>(in) + i);
            const OutT o0 = f(v.x), o1 = f(v.y), o2 = f(v.z), o3 = f(v.w);
            if constexpr (sizeof(OutT) == 4) {
                float4 o;
                o.x = o0; o.y = o1; o.z = o2; o.w = o3;
                __stcs(reinterpret_cast<float4 *>(out) + i, o);
            } else {
                uchar4 o;
                o.x = o0; o.y = o1; o.z = o2; o.w = o3;
                reinterpret_cast<uchar4 *>(out)[i] = o;
            }
        }
        for (size_t j = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < len; j += stride)
            out[j] = f(in[j]);
    } else {
        for (; i < len; i += stride) out[i] = f(in[i]);
    }
}

template <typename OutT, typename F>
ab_status launch_map1(ab_ctx *ctx, const float *in, OutT *out, size_t len, F f, cudaStream_t s) {
    if (len == 0) return AB_OK;
    size_t blocks = (len / 4 + PW_THREADS - 1) / PW_THREADS + 1;
    const size_t cap = (size_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    map1_kernel<OutT, F><<<(unsigned)blocks, PW_THREADS, 0, s>>>(in, out, len, f);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

StfTx make_tx(const ab_stf_params *p, const ab_image_stats *st) {
    StfTx t;
    const double range = std::fmax(st->max - st->min, 1e-30);
    const double clip_range = std::fmax(p->highlight - p->shadow, 1e-15);
    t.inv_range = 1.0 / range;
    t.dmin = st->min;
    t.shadow = p->shadow;
    t.inv_clip = 1.0 / clip_range;
    t.midtone = p->midtone;
    t.clip_range = clip_range;
    return t;
}

struct StfU8 {
    StfTx t;
    __device__ unsigned char operator()(float v) const {
        if (!valid_px(v)) return 0;
        // (x * 255).round().clamp(0, 255) as u8 -- round() is half-away-from-zero
        const double y = clampd(round(stf_apply(t, (double)v) * 255.0), 0.0, 255.0);
        return (y == y) ? (unsigned char)y : (unsigned char)0;
    }
};
struct StfF32 {
    StfTx t;
    __device__ float operator()(float v) const {
        return valid_px(v) ? (float)stf_apply(t, (double)v) : 0.0f;
    }
};
struct StfF32RgbVariant {  // compose/rgb.rs:191-207
    StfTx t;
    __device__ float operator()(float v) const {
        if (!(fabsf(v) < CUDART_INF_F) || v <= 1e-7f) return 0.0f;
        const double norm = ((double)v - t.dmin) * t.inv_range;
        const double clipped = clampd((norm - t.shadow) / t.clip_range, 0.0, 1.0);
        if (clipped <= 0.0) return 0.0f;
        if (clipped >= 1.0) return 1.0f;
        const double m = t.midtone;
        return (float)((m - 1.0) * clipped / ((2.0 * m - 1.0) * clipped - m));
    }
};

// Rust f32::asinh: (ax + ax / (hypot(1, 1/ax) + 1/ax)).ln_1p().copysign(x)
__device__ __forceinline__ float rust_asinhf(float x) {
    const float ax = fabsf(x);
    const float ix = 1.0f / ax;
    return copysignf(log1pf(ax + (ax / (hypotf(1.0f, ix) + ix))), x);
}
float rust_asinhf_host(float x) {
    const float ax = std::fabs(x);
    const float ix = 1.0f / ax;
    return std::copysign(std::log1p(ax + (ax / (std::hypot(1.0f, ix) + ix))), x);
}
struct Arcsinh {  // stretch.rs:28-44
    float dmin, inv_range, factor, inv_denom, gamma;
    bool apply_gamma;
    __device__ float operator()(float val) const {
        if (!(fabsf(val) < CUDART_INF_F)) return 0.0f;
        float norm = (val - dmin) * inv_range;
        norm = norm < 0.0f ? 0.0f : norm;
        norm = norm > 1.0f ? 1.0f : norm;
        const float stretched = rust_asinhf(norm * factor) * inv_denom;
        return apply_gamma ? powf(stretched, gamma) : stretched;
    }
};
struct Scale {
    float factor;
    __device__ float operator()(float v) const { return v * factor; }
};
struct Levels {  // curves.rs:25-50
    double black, inv_range, inv_gamma;
    __device__ float operator()(float v) const {
        if (!(fabsf(v) < CUDART_INF_F) || v < 0.0f) return 0.0f;
        const double norm = clampd(((double)v - black) * inv_range, 0.0, 1.0);
        return (float)pow(norm, inv_gamma);
    }
};
struct CurveLut {  // curves.rs:104-109,186-198; lut in global memory (16 KiB, L1/L2 resident)
    const float *lut;
    __device__ float operator()(float v) const {
        if (!(fabsf(v) < CUDART_INF_F) || v < 0.0f) return 0.0f;
        float c = v > 1.0f ? 1.0f : v;  // v >= 0 here
        unsigned idx = (unsigned)(c * 4095.0f);  // truncating, like `as usize`
        idx = idx > 4095u ? 4095u : idx;
        return __ldg(lut + idx);
    }
};

// channel_blend.rs:36-63 -- weights in constant-like kernel params, sequential mul-then-add
struct BlendParams {
    const float *ch[AB_MAX_BLEND_WEIGHTS];
    size_t len[AB_MAX_BLEND_WEIGHTS];
    float rw[AB_MAX_BLEND_WEIGHTS], gw[AB_MAX_BLEND_WEIGHTS], bw[AB_MAX_BLEND_WEIGHTS];
    int n_w;
};
__global__ void __launch_bounds__(PW_THREADS)
blend_kernel(const __grid_constant__ BlendParams bp, float *__restrict__ r, float *__restrict__ g,
             float *__restrict__ b, size_t npix, bool vec_ok) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n4 = vec_ok ? npix / 4 : 0;
    for (size_t i = tid; i < n4; i += stride) {
        float4 rv = make_float4(0.f, 0.f, 0.f, 0.f), gv = rv, bv = rv;
        for (int k = 0; k < bp.n_w; ++k) {
            if (i * 4 + 3 < bp.len[k]) {
                const float4 v = __ldcs(reinterpret_cast<const float4 *>(bp.ch[k]) + i);
                const float rw = bp.rw[k], gw = bp.gw[k], bw = bp.bw[k];
                rv.x += v.x * rw; rv.y += v.y * rw; rv.z += v.z * rw; rv.w += v.w * rw;
                gv.x += v.x * gw; gv.y += v.y * gw; gv.z += v.z * gw; gv.w += v.w * gw;
                bv.x += v.x * bw; bv.y += v.y * bw; bv.z += v.z * bw; bv.w += v.w * bw;
            } else {
                const float *c = bp.ch[k];
                const size_t L = bp.len[k];
                const float rw = bp.rw[k], gw = bp.gw[k], bw = bp.bw[k];
                float *rr = &rv.x, *gg = &gv.x, *bb = &bv.x;
                for (int q = 0; q < 4; ++q)
                    if (i * 4 + q < L) {
                        const float v = c[i * 4 + q];
                        rr[q] += v * rw; gg[q] += v * gw; bb[q] += v * bw;
                    }
            }
        }
        __stcs(reinterpret_cast<float4 *>(r) + i, rv);
        __stcs(reinterpret_cast<float4 *>(g) + i, gv);
        __stcs(reinterpret_cast<float4 *>(b) + i, bv);
    }
    for (size_t i = n4 * 4 + tid; i < npix; i += stride) {
        float rv = 0.f, gv = 0.f, bv = 0.f;
        for (int k = 0; k < bp.n_w; ++k)
            if (i < bp.len[k]) {
                const float v = bp.ch[k][i];
                rv += v * bp.rw[k];
                gv += v * bp.gw[k];
                bv += v * bp.bw[k];
            }
        r[i] = rv;
        g[i] = gv;
        b[i] = bv;
    }
}

// scnr.rs:36-52
__device__ __forceinline__ void scnr_px(float &rv, float &gv, float &bv, int method, float amount,
                                        bool preserve) {
    const float LUM_R = 0.2126f, LUM_G = 0.7152f, LUM_B = 0.0722f;
    const float INV_RB_WEIGHT = 1.0f / (LUM_R + LUM_B);
    const float limit = method == 0 ? (rv + bv) * 0.5f : fmaxf(rv, bv);
    const float g_corrected = fminf(gv, limit);
    const float g_new = gv + amount * (g_corrected - gv);
    const float delta_g = gv - g_new;
    if (preserve && delta_g > 1e-10f && rv <= 1.0f && bv <= 1.0f) {
        const float lum_lost = LUM_G * delta_g;
        const float boost = lum_lost * INV_RB_WEIGHT;
        const float rn = rv + boost, bn = bv + boost;
        rv = rn > 1.0f ? 1.0f : rn;
        bv = bn > 1.0f ? 1.0f : bn;
    }
    gv = g_new;
}
__global__ void __launch_bounds__(PW_THREADS)
scnr_kernel(float *__restrict__ r, float *__restrict__ g, float *__restrict__ b, size_t len,
            int method, float amount, bool preserve, bool vec_ok) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n4 = vec_ok ? len / 4 : 0;
    for (size_t i = tid; i < n4; i += stride) {
        float4 rv = reinterpret_cast<float4 *>(r)[i];
        float4 gv = reinterpret_cast<float4 *>(g)[i];
        float4 bv = reinterpret_cast<float4 *>(b)[i];
        scnr_px(rv.x, gv.x, bv.x, method, amount, preserve);
        scnr_px(rv.y, gv.y, bv.y, method, amount, preserve);
        scnr_px(rv.z, gv.z, bv.z, method, amount, preserve);
        scnr_px(rv.w, gv.w, bv.w, method, amount, preserve);
        if (preserve) {
            reinterpret_cast<float4 *>(r)[i] = rv;
            reinterpret_cast<float4 *>(b)[i] = bv;
        }
        reinterpret_cast<float4 *>(g)[i] = gv;
    }
    for (size_t i = n4 * 4 + tid; i < len; i += stride) {
        float rv = r[i], gv = g[i], bv = b[i];
        scnr_px(rv, gv, bv, method, amount, preserve);
        if (preserve) {
            r[i] = rv;
            b[i] = bv;
        }
        g[i] = gv;
    }
}

// ---- fused tone chain (apply_tone_composite_cmd, cmd/processing/curves.rs:111-151): per channel
// apply_stf_f32 -> apply_levels (unless identity) -> apply_curve (when curves are applied), then
// apply_scnr_inplace across the channels; one read and one write of the three planes instead of
// four sweeps.  Every stage is the same device function the standalone operators use.
struct ToneChain {
    StfTx stf[3];
    Levels lv[3];
    bool lv_on[3];
    const float *lut[3];
    bool curves, scnr, preserve;
    int method;
    float amount;
};
__device__ __forceinline__ float tone_px(const ToneChain &c, int ch, float v) {
    v = StfF32{c.stf[ch]}(v);
    if (c.lv_on[ch]) v = c.lv[ch](v);
    if (c.curves) v = CurveLut{c.lut[ch]}(v);
    return v;
}
__global__ void __launch_bounds__(PW_THREADS)
tone_chain_kernel(const float *r, const float *g, const float *b, size_t len,
                  const __grid_constant__ ToneChain c, float *out_r, float *out_g, float *out_b,
                  bool vec_ok) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n4 = vec_ok ? len / 4 : 0;
    for (size_t i = tid; i < n4; i += stride) {
        float4 rv = __ldcs(reinterpret_cast<const float4 *>(r) + i);
        float4 gv = __ldcs(reinterpret_cast<const float4 *>(g) + i);
        float4 bv = __ldcs(reinterpret_cast<const float4 *>(b) + i);
        float *rr = &rv.x, *gg = &gv.x, *bb = &bv.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            rr[q] = tone_px(c, 0, rr[q]);
            gg[q] = tone_px(c, 1, gg[q]);
            bb[q] = tone_px(c, 2, bb[q]);
            if (c.scnr) scnr_px(rr[q], gg[q], bb[q], c.method, c.amount, c.preserve);
        }
        __stcs(reinterpret_cast<float4 *>(out_r) + i, rv);
        __stcs(reinterpret_cast<float4 *>(out_g) + i, gv);
        __stcs(reinterpret_cast<float4 *>(out_b) + i, bv);
    }
    for (size_t i = n4 * 4 + tid; i < len; i += stride) {
        float rv = tone_px(c, 0, r[i]), gv = tone_px(c, 1, g[i]), bv = tone_px(c, 2, b[i]);
        if (c.scnr) scnr_px(rv, gv, bv, c.method, c.amount, c.preserve);
        out_r[i] = rv;
        out_g[i] = gv;
        out_b[i] = bv;
    }
}

unsigned sweep_blocks(ab_ctx *ctx, size_t len) {
    size_t blocks = (len / 4 + PW_THREADS - 1) / PW_THREADS + 1;
    const size_t cap = (size_t)ctx->sm_count * 8;
    return (unsigned)(blocks > cap ? cap : blocks);
}

// Host-buffer helper: copy `n_in` planes in, run `body(d_in[], d_out[])` on the context stream,
// copy `n_out` planes back, synchronize.  Scratch slot 0 is carved into 256-byte aligned planes.
template <typename Body>
ab_status with_host_planes(ab_ctx *ctx, const void *const *h_in, const size_t *in_bytes, int n_in,
                           void *const *h_out, const size_t *out_bytes, int n_out, Body body) {
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    size_t total = 0;
    std::vector<size_t> off;
    for (int i = 0; i < n_in; ++i) {
        off.push_back(total);
        total += (in_bytes[i] + 255) / 256 * 256;
    }
    for (int i = 0; i < n_out; ++i) {
        off.push_back(total);
        total += (out_bytes[i] + 255) / 256 * 256;
    }
    ab_status st = ab_scratch_reserve(ctx, 0, total ? total : 256);
    if (st != AB_OK) return st;
    char *base = static_cast<char *>(ctx->d_scratch[0]);
    std::vector<void *> d_in(n_in), d_out(n_out);
    for (int i = 0; i < n_in; ++i) {
        d_in[i] = base + off[i];
        if (in_bytes[i])
            AB_CUDA(ctx, cudaMemcpyAsync(d_in[i], h_in[i], in_bytes[i], cudaMemcpyHostToDevice,
                                         ctx->stream));
    }
    for (int i = 0; i < n_out; ++i) d_out[i] = base + off[n_in + i];
    st = body(d_in.data(), d_out.data(), ctx->stream);
    if (st != AB_OK) return st;
    for (int i = 0; i < n_out; ++i)
        if (out_bytes[i])
            AB_CUDA(ctx, cudaMemcpyAsync(h_out[i], d_out[i], out_bytes[i], cudaMemcpyDeviceToHost,
                                         ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------ STF
ab_status ab_apply_stf_u8_dev(ab_ctx *ctx, const float *d_data, size_t len,
                              const ab_stf_params *p, const ab_image_stats *st, uint8_t *d_out,
                              void *stream) {
    if (!ctx || !p || !st || (len && (!d_data || !d_out))) return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    return launch_map1<unsigned char>(ctx, d_data, d_out, len, StfU8{make_tx(p, st)},
                                      (cudaStream_t)stream);
}
ab_status ab_apply_stf_f32_dev(ab_ctx *ctx, const float *d_data, size_t len,
                               const ab_stf_params *p, const ab_image_stats *st, int variant,
                               float *d_out, void *stream) {
    if (!ctx || !p || !st || (len && (!d_data || !d_out))) return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    if (variant == 1)
        return launch_map1<float>(ctx, d_data, d_out, len, StfF32RgbVariant{make_tx(p, st)},
                                  (cudaStream_t)stream);
    return launch_map1<float>(ctx, d_data, d_out, len, StfF32{make_tx(p, st)},
                              (cudaStream_t)stream);
}
ab_status ab_apply_stf_u8(ab_ctx *ctx, const float *data, size_t len, const ab_stf_params *p,
                          const ab_image_stats *st, uint8_t *out) {
    if (!ctx || !p || !st || (len && (!data || !out))) return AB_ERR_INVALID;
    const void *in[1] = {data};
    void *o[1] = {out};
    const size_t ib[1] = {len * 4}, ob[1] = {len};
    return with_host_planes(ctx, in, ib, 1, o, ob, 1, [&](void **di, void **dout, cudaStream_t s) {
        return launch_map1<unsigned char>(ctx, (const float *)di[0], (unsigned char *)dout[0], len,
                                          StfU8{make_tx(p, st)}, s);
    });
}
ab_status ab_apply_stf_f32(ab_ctx *ctx, const float *data, size_t len, const ab_stf_params *p,
                           const ab_image_stats *st, int variant, float *out) {
    if (!ctx || !p || !st || (len && (!data || !out))) return AB_ERR_INVALID;
    const void *in[1] = {data};
    void *o[1] = {out};
    const size_t ib[1] = {len * 4}, ob[1] = {len * 4};
    return with_host_planes(ctx, in, ib, 1, o, ob, 1, [&](void **di, void **dout, cudaStream_t s) {
        if (variant == 1)
            return launch_map1<float>(ctx, (const float *)di[0], (float *)dout[0], len,
                                      StfF32RgbVariant{make_tx(p, st)}, s);
        return launch_map1<float>(ctx, (const float *)di[0], (float *)dout[0], len,
                                  StfF32{make_tx(p, st)}, s);
    });
}

// ------------------------------------------------------------------------------------ arcsinh
// Returns 1 when the reference short-circuits (copy / zeros) and fills `mode`: 1 copy, 2 zeros.
static int arcsinh_setup(float dmin, float dmax, float factor, float gamma, Arcsinh *a) {
    if (std::fabs(factor) < 1e-10f) return 1;  // stretch.rs:17-19: clone
    const float range = dmax - dmin;
    if (range < 1e-10f) return 2;  // stretch.rs:21-24: zeros (NaN range falls through like Rust)
    a->dmin = dmin;
    a->inv_range = 1.0f / range;
    a->factor = factor;
    a->inv_denom = 1.0f / rust_asinhf_host(factor);
    a->gamma = gamma;
    a->apply_gamma = std::fabs(gamma - 1.0f) > 1e-6f;
    return 0;
}
ab_status ab_arcsinh_stretch_dev(ab_ctx *ctx, const float *d_data, size_t len, float dmin,
                                 float dmax, float factor, float gamma, float *d_out,
                                 void *stream) {
    if (!ctx || (len && (!d_data || !d_out))) return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    Arcsinh a;
    const int mode = arcsinh_setup(dmin, dmax, factor, gamma, &a);
    cudaStream_t s = (cudaStream_t)stream;
    if (mode == 1) {
        AB_CUDA(ctx, cudaMemcpyAsync(d_out, d_data, len * 4, cudaMemcpyDeviceToDevice, s));
        return AB_OK;
    }
    if (mode == 2) {
        AB_CUDA(ctx, cudaMemsetAsync(d_out, 0, len * 4, s));
        return AB_OK;
    }
    return launch_map1<float>(ctx, d_data, d_out, len, a, s);
}
ab_status ab_arcsinh_stretch(ab_ctx *ctx, const float *data, size_t len, float dmin, float dmax,
                             float factor, float gamma, float *out) {
    if (!ctx || (len && (!data || !out))) return AB_ERR_INVALID;
    Arcsinh a;
    const int mode = arcsinh_setup(dmin, dmax, factor, gamma, &a);
    if (mode == 1) {
        std::memcpy(out, data, len * 4);
        return AB_OK;
    }
    if (mode == 2) {
        std::memset(out, 0, len * 4);
        return AB_OK;
    }
    const void *in[1] = {data};
    void *o[1] = {out};
    const size_t ib[1] = {len * 4};
    return with_host_planes(ctx, in, ib, 1, o, ib, 1, [&](void **di, void **dout, cudaStream_t s) {
        return launch_map1<float>(ctx, (const float *)di[0], (float *)dout[0], len, a, s);
    });
}

// ------------------------------------------------------------------------------------ scale
ab_status ab_scale_dev(ab_ctx *ctx, const float *d_data, size_t len, float factor, float *d_out,
                       void *stream) {
    if (!ctx || (len && (!d_data || !d_out))) return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    return launch_map1<float>(ctx, d_data, d_out, len, Scale{factor}, (cudaStream_t)stream);
}
ab_status ab_scale(ab_ctx *ctx, const float *data, size_t len, float factor, float *out) {
    if (!ctx || (len && (!data || !out))) return AB_ERR_INVALID;
    const void *in[1] = {data};
    void *o[1] = {out};
    const size_t ib[1] = {len * 4};
    return with_host_planes(ctx, in, ib, 1, o, ib, 1, [&](void **di, void **dout, cudaStream_t s) {
        return launch_map1<float>(ctx, (const float *)di[0], (float *)dout[0], len, Scale{factor}, s);
    });
}

// ------------------------------------------------------------------------------------ levels
static bool levels_identity(const ab_levels_params *p) {  // curves.rs:17-23
    return std::fabs(p->black) < 1e-7 && std::fabs(p->gamma - 1.0) < 1e-7 &&
           std::fabs(p->white - 1.0) < 1e-7;
}
static Levels levels_make(const ab_levels_params *p) {
    Levels l;
    const double range = std::fmax(p->white - p->black, 1e-15);
    l.black = p->black;
    l.inv_range = 1.0 / range;
    double gm = p->gamma;
    gm = gm < 0.01 ? 0.01 : (gm > 10.0 ? 10.0 : gm);
    l.inv_gamma = 1.0 / gm;
    return l;
}
ab_status ab_apply_levels_dev(ab_ctx *ctx, const float *d_data, size_t len,
                              const ab_levels_params *p, float *d_out, void *stream) {
    if (!ctx || !p || (len && (!d_data || !d_out))) return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (levels_identity(p)) {
        AB_CUDA(ctx, cudaMemcpyAsync(d_out, d_data, len * 4, cudaMemcpyDeviceToDevice, s));
        return AB_OK;
    }
    return launch_map1<float>(ctx, d_data, d_out, len, levels_make(p), s);
}
ab_status ab_apply_levels(ab_ctx *ctx, const float *data, size_t len, const ab_levels_params *p,
                          float *out) {
    if (!ctx || !p || (len && (!data || !out))) return AB_ERR_INVALID;
    if (levels_identity(p)) {
        std::memcpy(out, data, len * 4);
        return AB_OK;
    }
    const void *in[1] = {data};
    void *o[1] = {out};
    const size_t ib[1] = {len * 4};
    return with_host_planes(ctx, in, ib, 1, o, ib, 1, [&](void **di, void **dout, cudaStream_t s) {
        return launch_map1<float>(ctx, (const float *)di[0], (float *)dout[0], len, levels_make(p), s);
    });
}

// ------------------------------------------------------------------------------------ curves
// SplineLut::from_points, curves.rs:70-95 with fritsch_carlson_tangents :113-151 and
// hermite_eval :153-184.  Pure host scalar math (4096 evaluations), like the reference.
ab_status ab_spline_lut(const double *points_xy, size_t n_pts, float *lut) {
    if (!lut || (n_pts && !points_xy)) return AB_ERR_INVALID;
    std::vector<std::pair<double, double>> pts;
    for (size_t i = 0; i < n_pts; ++i) pts.push_back({points_xy[2 * i], points_xy[2 * i + 1]});
    std::stable_sort(pts.begin(), pts.end(),
                     [](const std::pair<double, double> &a, const std::pair<double, double> &b) {
                         return a.first < b.first;
                     });
    {
        std::vector<std::pair<double, double>> d;
        for (const auto &p : pts) {
            if (!d.empty() && std::fabs(p.first - d.back().first) < 1e-9) continue;
            d.push_back(p);
        }
        pts.swap(d);
    }
    if (pts.empty() || pts[0].first > 1e-6) pts.insert(pts.begin(), {0.0, 0.0});
    if (pts.back().first < 1.0 - 1e-6) pts.push_back({1.0, 1.0});
    const size_t n = pts.size();
    std::vector<double> m(n, 0.0);
    if (n == 2) {
        const double slope =
            (pts[1].second - pts[0].second) / std::fmax(pts[1].first - pts[0].first, 1e-15);
        m[0] = m[1] = slope;
    } else if (n > 2) {
        std::vector<double> slopes;
        for (size_t i = 0; i + 1 < n; ++i) {
            const double dx = std::fmax(pts[i + 1].first - pts[i].first, 1e-15);
            slopes.push_back((pts[i + 1].second - pts[i].second) / dx);
        }
        auto signum = [](double x) { return std::isnan(x) ? x : (std::signbit(x) ? -1.0 : 1.0); };
        m[0] = slopes[0];
        m[n - 1] = slopes[n - 2];
        for (size_t i = 1; i + 1 < n; ++i)
            m[i] = signum(slopes[i - 1]) != signum(slopes[i]) ? 0.0
                                                              : (slopes[i - 1] + slopes[i]) * 0.5;
        for (size_t i = 0; i + 1 < n; ++i) {
            if (std::fabs(slopes[i]) < 1e-15) {
                m[i] = 0.0;
                m[i + 1] = 0.0;
                continue;
            }
            const double alpha = m[i] / slopes[i], beta = m[i + 1] / slopes[i];
            const double tau = alpha * alpha + beta * beta;
            if (tau > 9.0) {
                const double sc = 3.0 / std::sqrt(tau);
                m[i] = sc * alpha * slopes[i];
                m[i + 1] = sc * beta * slopes[i];
            }
        }
    }
    for (int i = 0; i < AB_CURVE_LUT_SIZE; ++i) {
        const double x = (double)i / 4095.0;
        double y;
        if (x <= pts[0].first) {
            y = pts[0].second;
        } else if (x >= pts[n - 1].first) {
            y = pts[n - 1].second;
        } else {
            size_t seg = 0;
            for (size_t k = 1; k < n; ++k)
                if (x < pts[k].first) {
                    seg = k - 1;
                    break;
                }
            const double dx = std::fmax(pts[seg + 1].first - pts[seg].first, 1e-15);
            const double t = (x - pts[seg].first) / dx;
            const double t2 = t * t, t3 = t2 * t;
            const double h00 = 2.0 * t3 - 3.0 * t2 + 1.0, h10 = t3 - 2.0 * t2 + t;
            const double h01 = -2.0 * t3 + 3.0 * t2, h11 = t3 - t2;
            y = h00 * pts[seg].second + h10 * dx * m[seg] + h01 * pts[seg + 1].second +
                h11 * dx * m[seg + 1];
        }
        y = y < 0.0 ? 0.0 : (y > 1.0 ? 1.0 : y);
        lut[i] = (float)y;
    }
    return AB_OK;
}
ab_status ab_apply_curve_dev(ab_ctx *ctx, const float *d_data, size_t len, const float *d_lut,
                             float *d_out, void *stream) {
    if (!ctx || !d_lut || (len && (!d_data || !d_out))) return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    return launch_map1<float>(ctx, d_data, d_out, len, CurveLut{d_lut}, (cudaStream_t)stream);
}
ab_status ab_apply_curve(ab_ctx *ctx, const float *data, size_t len, const float *lut, float *out) {
    if (!ctx || !lut || (len && (!data || !out))) return AB_ERR_INVALID;
    const void *in[2] = {data, lut};
    void *o[1] = {out};
    const size_t ib[2] = {len * 4, AB_CURVE_LUT_SIZE * 4}, ob[1] = {len * 4};
    return with_host_planes(ctx, in, ib, 2, o, ob, 1, [&](void **di, void **dout, cudaStream_t s) {
        return launch_map1<float>(ctx, (const float *)di[0], (float *)dout[0], len,
                                  CurveLut{(const float *)di[1]}, s);
    });
}

// ------------------------------------------------------------------------------------ blend
static ab_status blend_setup(ab_ctx *ctx, const float *const *ch, const size_t *lens, size_t n_ch,
                             const ab_blend_weight *w, size_t n_w, BlendParams *bp) {
    bp->n_w = 0;
    for (size_t k = 0; k < n_w; ++k) {
        if (w[k].channel_idx >= n_ch) continue;  // channel_blend.rs:21-23
        if (bp->n_w >= AB_MAX_BLEND_WEIGHTS)
            return ab_fail(ctx, AB_ERR_UNSUPPORTED, "more than %d blend weights",
                           AB_MAX_BLEND_WEIGHTS);
        const int i = bp->n_w++;
        bp->ch[i] = ch[w[k].channel_idx];
        bp->len[i] = lens[w[k].channel_idx];
        bp->rw[i] = (float)w[k].r_weight;
        bp->gw[i] = (float)w[k].g_weight;
        bp->bw[i] = (float)w[k].b_weight;
    }
    return AB_OK;
}
static ab_status blend_launch(ab_ctx *ctx, const BlendParams &bp, float *r, float *g, float *b,
                              size_t npix, cudaStream_t s) {
    if (npix == 0) return AB_OK;
    bool vec_ok = ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(g) |
                    reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    for (int k = 0; k < bp.n_w; ++k) vec_ok = vec_ok && ((reinterpret_cast<uintptr_t>(bp.ch[k]) & 15) == 0);
    blend_kernel<<<sweep_blocks(ctx, npix), PW_THREADS, 0, s>>>(bp, r, g, b, npix, vec_ok);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}
ab_status ab_blend_channels_dev(ab_ctx *ctx, const float *const *d_channels,
                                const size_t *channel_lens, size_t n_ch, const ab_blend_weight *w,
                                size_t n_w, size_t rows, size_t cols, float *d_r, float *d_g,
                                float *d_b, void *stream) {
    if (!ctx || (n_ch && (!d_channels || !channel_lens)) || (n_w && !w) || !d_r || !d_g || !d_b)
        return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    BlendParams bp;
    ab_status st = blend_setup(ctx, d_channels, channel_lens, n_ch, w, n_w, &bp);
    if (st != AB_OK) return st;
    return blend_launch(ctx, bp, d_r, d_g, d_b, rows * cols, (cudaStream_t)stream);
}
ab_status ab_blend_channels(ab_ctx *ctx, const float *const *channels, const size_t *channel_lens,
                            size_t n_ch, const ab_blend_weight *w, size_t n_w, size_t rows,
                            size_t cols, float *r, float *g, float *b) {
    if (!ctx || (n_ch && (!channels || !channel_lens)) || (n_w && !w) || !r || !g || !b)
        return AB_ERR_INVALID;
    if (n_ch > AB_MAX_BLEND_WEIGHTS)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "more than %d channels", AB_MAX_BLEND_WEIGHTS);
    const size_t npix = rows * cols;
    std::vector<const void *> in(n_ch);
    std::vector<size_t> ib(n_ch);
    for (size_t c = 0; c < n_ch; ++c) {
        in[c] = channels[c];
        ib[c] = channel_lens[c] * 4;
    }
    void *o[3] = {r, g, b};
    const size_t ob[3] = {npix * 4, npix * 4, npix * 4};
    return with_host_planes(ctx, in.data(), ib.data(), (int)n_ch, o, ob, 3,
                            [&](void **di, void **dout, cudaStream_t s) {
                                std::vector<const float *> dch(n_ch);
                                for (size_t c = 0; c < n_ch; ++c) dch[c] = (const float *)di[c];
                                BlendParams bp;
                                ab_status st = blend_setup(ctx, dch.data(), channel_lens, n_ch, w, n_w, &bp);
                                if (st != AB_OK) return st;
                                return blend_launch(ctx, bp, (float *)dout[0], (float *)dout[1],
                                                    (float *)dout[2], npix, s);
                            });
}

// ------------------------------------------------------------------------------------ SCNR
ab_status ab_apply_scnr_dev(ab_ctx *ctx, float *d_r, float *d_g, float *d_b, size_t len,
                            const ab_scnr_config *cfg, void *stream) {
    if (!ctx || !cfg || (len && (!d_r || !d_g || !d_b))) return AB_ERR_INVALID;
    float amount = cfg->amount;
    amount = amount < 0.0f ? 0.0f : (amount > 1.0f ? 1.0f : amount);
    if (amount < 1e-7f || len == 0) return AB_OK;  // scnr.rs:28-31 (NaN amount falls through)
    DeviceGuard g(ctx->device);
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(d_r) | reinterpret_cast<uintptr_t>(d_g) |
                          reinterpret_cast<uintptr_t>(d_b)) & 15) == 0;
    scnr_kernel<<<sweep_blocks(ctx, len), PW_THREADS, 0, (cudaStream_t)stream>>>(
        d_r, d_g, d_b, len, cfg->method, amount, cfg->preserve_luminance != 0, vec_ok);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}
ab_status ab_apply_scnr(ab_ctx *ctx, float *r, float *g, float *b, size_t len,
                        const ab_scnr_config *cfg) {
    if (!ctx || !cfg || (len && (!r || !g || !b))) return AB_ERR_INVALID;
    float amount = cfg->amount;
    amount = amount < 0.0f ? 0.0f : (amount > 1.0f ? 1.0f : amount);
    if (amount < 1e-7f || len == 0) return AB_OK;
    const void *in[3] = {r, g, b};
    void *o[3] = {r, g, b};
    const size_t ib[3] = {len * 4, len * 4, len * 4};
    // in place on the device copies: outputs alias inputs
    CtxLock lk(ctx->mu);
    DeviceGuard gd(ctx->device);
    const size_t plane = (len * 4 + 255) / 256 * 256;
    ab_status st = ab_scratch_reserve(ctx, 0, plane * 3);
    if (st != AB_OK) return st;
    char *base = static_cast<char *>(ctx->d_scratch[0]);
    float *d[3];
    for (int i = 0; i < 3; ++i) {
        d[i] = reinterpret_cast<float *>(base + i * plane);
        AB_CUDA(ctx, cudaMemcpyAsync(d[i], in[i], ib[i], cudaMemcpyHostToDevice, ctx->stream));
    }
    const bool preserve = cfg->preserve_luminance != 0;
    scnr_kernel<<<sweep_blocks(ctx, len), PW_THREADS, 0, ctx->stream>>>(
        d[0], d[1], d[2], len, cfg->method, amount, preserve, true);
    AB_LAUNCH_CHECK(ctx);
    for (int i = 0; i < 3; ++i)
        if (i == 1 || preserve)
            AB_CUDA(ctx, cudaMemcpyAsync(o[i], d[i], ib[i], cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}

// ------------------------------------------------------------------------------------ tone chain
static bool tone_setup(const ab_tone_chain *t, const float *const *d_lut, ToneChain *c) {
    for (int ch = 0; ch < 3; ++ch) {
        c->stf[ch] = make_tx(&t->stf[ch], &t->stats[ch]);
        c->lv_on[ch] = !levels_identity(&t->levels[ch]);
        c->lv[ch] = levels_make(&t->levels[ch]);
        c->lut[ch] = d_lut ? d_lut[ch] : nullptr;
    }
    c->curves = d_lut != nullptr;
    float amount = t->scnr.amount;
    amount = amount < 0.0f ? 0.0f : (amount > 1.0f ? 1.0f : amount);
    // curves.rs:143-148: applied when cfg.amount > 1e-7 (scnr.rs:28-31 clamps, then re-checks)
    c->scnr = t->apply_scnr != 0 && t->scnr.amount > 1e-7f && !(amount < 1e-7f);
    c->amount = amount;
    c->method = t->scnr.method;
    c->preserve = t->scnr.preserve_luminance != 0;
    return true;
}
ab_status ab_tone_composite_dev(ab_ctx *ctx, const float *d_r, const float *d_g, const float *d_b,
                                size_t len, const ab_tone_chain *chain, float *d_out_r,
                                float *d_out_g, float *d_out_b, void *stream) {
    if (!ctx || !chain || (len && (!d_r || !d_g || !d_b || !d_out_r || !d_out_g || !d_out_b)))
        return AB_ERR_INVALID;
    const bool any_lut = chain->lut[0] || chain->lut[1] || chain->lut[2];
    if (any_lut && !(chain->lut[0] && chain->lut[1] && chain->lut[2]))
        return ab_fail(ctx, AB_ERR_INVALID, "tone chain: give all three curve LUTs or none");
    if (len == 0) return AB_OK;
    DeviceGuard g(ctx->device);
    ToneChain c;
    tone_setup(chain, any_lut ? chain->lut : nullptr, &c);
    uintptr_t al = 0;
    for (const void *q : {(const void *)d_r, (const void *)d_g, (const void *)d_b,
                          (const void *)d_out_r, (const void *)d_out_g, (const void *)d_out_b})
        al |= reinterpret_cast<uintptr_t>(q);
    tone_chain_kernel<<<sweep_blocks(ctx, len), PW_THREADS, 0, (cudaStream_t)stream>>>(
        d_r, d_g, d_b, len, c, d_out_r, d_out_g, d_out_b, (al & 15) == 0);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}
ab_status ab_tone_composite(ab_ctx *ctx, const float *r, const float *g, const float *b, size_t len,
                            const ab_tone_chain *chain, float *out_r, float *out_g, float *out_b) {
    if (!ctx || !chain || (len && (!r || !g || !b || !out_r || !out_g || !out_b)))
        return AB_ERR_INVALID;
    const bool any_lut = chain->lut[0] || chain->lut[1] || chain->lut[2];
    if (any_lut && !(chain->lut[0] && chain->lut[1] && chain->lut[2]))
        return ab_fail(ctx, AB_ERR_INVALID, "tone chain: give all three curve LUTs or none");
    if (len == 0) return AB_OK;
    const size_t lb = any_lut ? AB_CURVE_LUT_SIZE * 4 : 0;
    const void *in[6] = {r, g, b, chain->lut[0], chain->lut[1], chain->lut[2]};
    void *o[3] = {out_r, out_g, out_b};
    const size_t ib[6] = {len * 4, len * 4, len * 4, lb, lb, lb}, ob[3] = {len * 4, len * 4, len * 4};
    return with_host_planes(ctx, in, ib, 6, o, ob, 3, [&](void **di, void **dout, cudaStream_t s) {
        ToneChain c;
        const float *luts[3] = {(const float *)di[3], (const float *)di[4], (const float *)di[5]};
        tone_setup(chain, any_lut ? luts : nullptr, &c);
        tone_chain_kernel<<<sweep_blocks(ctx, len), PW_THREADS, 0, s>>>(
            (const float *)di[0], (const float *)di[1], (const float *)di[2], len, c,
            (float *)dout[0], (float *)dout[1], (float *)dout[2], true);
        AB_LAUNCH_CHECK(ctx);
        return AB_OK;
    });
}

}  // extern "C"
