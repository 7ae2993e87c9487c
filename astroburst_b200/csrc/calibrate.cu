// calibrate.cu -- calibration masters and frame calibration.
//
// Reference: src-tauri/src/core/stacking/calibration.rs
//   median_combine_row_major :85-123 (per-pixel upper median of the finite samples, empty -> 0),
//   the per-frame preprocessing of create_master_dark :160-186 (frame - bias) and
//   create_master_flat :199-229 (frame - bias, then - dark * 1.0),
//   the master-flat normalisation :231-252, calibrate_image :47-83,
//   subtract_bias / subtract_dark / divide_flat :15-45 (calibrate without the final clamp).
// The median combine is the chunked deep-stack kernel (stack_deep.cu) in its median-only mode:
// the same register sorting network, rank selection on the sorted chunks, any frame count <= 256.
#include <algorithm>

#include <math_constants.h>

#include "ab_internal.h"

namespace {

constexpr int CAL_THREADS = 256;

// calibrate_image, calibration.rs:64-80 -- one expression per rounding, like the reference
__global__ void __launch_bounds__(CAL_THREADS)
calibrate_kernel(const float *__restrict__ raw, size_t len, const float *__restrict__ bias,
                 const float *__restrict__ dark, const float *__restrict__ flat, float dark_ratio,
                 bool clamp_negative, float *__restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        float v = __ldcs(raw + i);
        if (bias) v -= __ldcs(bias + i);
        if (dark) v -= __ldcs(dark + i) * dark_ratio;
        if (flat) {
            const float fv = __ldcs(flat + i);
            if (fabsf(fv) < CUDART_INF_F && fabsf(fv) > 1e-4f) v /= fv;
        }
        if (clamp_negative && v < 0.0f) v = 0.0f;  // (NaN stays NaN: `v < 0.0` is false)
        __stcs(out + i, v);
    }
}

// master-flat normalisation, calibration.rs:231-252.  The reference sums the finite positive
// pixels sequentially in f64; here: per-thread, per-block and final sums in a fixed order
// (deterministic; equal to the reference's mean to ~1e-16 before the cast to f32).
struct FlatPartial {
    double sum;
    unsigned long long count;
};
constexpr int FLAT_BLOCKS = 592;  // 4 per SM

__global__ void __launch_bounds__(CAL_THREADS)
flat_sum_kernel(const float *__restrict__ d, size_t len, FlatPartial *__restrict__ partials) {
    double sum = 0.0;
    unsigned long long cnt = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const float v = d[i];
        if (fabsf(v) < CUDART_INF_F && v > 0.0f) {
            sum += (double)v;
            ++cnt;
        }
    }
    __shared__ FlatPartial sh[CAL_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = {sum, cnt};
    __syncthreads();
    if (threadIdx.x == 0) {
        FlatPartial r = sh[0];
        for (int w = 1; w < CAL_THREADS / 32; ++w) {
            r.sum += sh[w].sum;
            r.count += sh[w].count;
        }
        partials[blockIdx.x] = r;
    }
}
// one thread: total, then inv_mean as calibration.rs:241-242; factor[1] = 1 when count > 0
__global__ void flat_factor_kernel(const FlatPartial *__restrict__ partials, int n, float *factor) {
    double sum = 0.0;
    unsigned long long cnt = 0;
    for (int i = 0; i < n; ++i) {
        sum += partials[i].sum;
        cnt += partials[i].count;
    }
    float inv_mean = 1.0f;
    if (cnt > 0) {
        const double mean = sum / (double)cnt;
        inv_mean = fabs(mean) > 1e-10 ? 1.0f / (float)mean : 1.0f;
    }
    factor[0] = inv_mean;
    factor[1] = cnt > 0 ? 1.0f : 0.0f;
}
__global__ void __launch_bounds__(CAL_THREADS)
flat_scale_kernel(float *__restrict__ d, size_t len, const float *__restrict__ factor) {
    if (factor[1] == 0.0f) return;  // no finite positive pixel: left untouched (calibration.rs:239)
    const float inv_mean = factor[0];
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const float v = d[i];
        d[i] = (fabsf(v) < CUDART_INF_F && v > 0.0f) ? v * inv_mean : 1.0f;
    }
}

unsigned cal_blocks(ab_ctx *ctx, size_t len) {
    size_t b = (len + CAL_THREADS - 1) / CAL_THREADS;
    const size_t cap = (size_t)ctx->sm_count * 16;
    return (unsigned)std::max<size_t>(1, std::min(b, cap));
}

ab_status normalize_flat_on(ab_ctx *ctx, float *d, size_t len, cudaStream_t s) {
    // scratch in d_small (stats.cu owns [0, 640 KiB)): [704 KiB, +FLAT_BLOCKS partials) | factor
    char *base = static_cast<char *>(ctx->d_small) + (704 << 10);
    FlatPartial *partials = reinterpret_cast<FlatPartial *>(base);
    float *factor = reinterpret_cast<float *>(base + FLAT_BLOCKS * sizeof(FlatPartial));
    const int blocks = (int)std::min<size_t>(FLAT_BLOCKS, std::max<size_t>(1, (len + CAL_THREADS - 1) / CAL_THREADS));
    flat_sum_kernel<<<blocks, CAL_THREADS, 0, s>>>(d, len, partials);
    AB_LAUNCH_CHECK(ctx);
    flat_factor_kernel<<<1, 1, 0, s>>>(partials, blocks, factor);
    AB_LAUNCH_CHECK(ctx);
    flat_scale_kernel<<<cal_blocks(ctx, len), CAL_THREADS, 0, s>>>(d, len, factor);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

}  // namespace

extern "C" {

ab_status ab_median_combine_dev(ab_ctx *ctx, const float *const *d_frames, size_t n_frames,
                                size_t npix, const float *d_bias, const float *d_dark,
                                float dark_ratio, float *d_out, void *stream) {
    if (!ctx) return AB_ERR_INVALID;
    if (n_frames == 0) return ab_fail(ctx, AB_ERR_INVALID, "No frames provided");
    if (!d_frames || (npix && !d_out)) return ab_fail(ctx, AB_ERR_INVALID, "null pointer");
    if (n_frames > AB_MAX_FRAMES)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "n_frames %zu > AB_MAX_FRAMES %d", n_frames,
                       AB_MAX_FRAMES);
    if (npix == 0) return AB_OK;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    FrameTable ft;
    for (size_t f = 0; f < AB_MAX_FRAMES; ++f) ft.p[f] = f < n_frames ? d_frames[f] : nullptr;
    DeepPre pre;
    pre.bias = d_bias;
    pre.dark = d_dark;
    pre.dark_ratio = dark_ratio;
    pre.median_only = 1;
    // the kernel's rejected counter is unused in this mode; any valid device word will do
    unsigned long long *d_rej = reinterpret_cast<unsigned long long *>(
        static_cast<char *>(ctx->d_small) + (768 << 10));
    return ab_stack_launch_deep(ctx, ft, (int)n_frames, npix, 0.0f, 0.0f, 0u, pre, d_out, d_rej,
                                (cudaStream_t)stream);
}

ab_status ab_median_combine(ab_ctx *ctx, const float *const *frames, size_t n_frames, size_t rows,
                            size_t cols, const float *bias, const float *dark, float dark_ratio,
                            float *out) {
    if (!ctx) return AB_ERR_INVALID;
    if (n_frames == 0) return ab_fail(ctx, AB_ERR_INVALID, "No frames provided");
    if (!frames || !out) return ab_fail(ctx, AB_ERR_INVALID, "null pointer");
    if (n_frames > AB_MAX_FRAMES)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "n_frames %zu > AB_MAX_FRAMES %d", n_frames,
                       AB_MAX_FRAMES);
    for (size_t f = 0; f < n_frames; ++f)
        if (!frames[f]) return ab_fail(ctx, AB_ERR_INVALID, "null frame pointer %zu", f);
    const size_t npix = rows * cols;
    if (npix == 0) return AB_OK;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    // row bands of ~96 MiB of input, copies of band b+1 overlap the kernel of band b
    const size_t planes = n_frames + 1 + (bias ? 1 : 0) + (dark ? 1 : 0);
    const size_t target_px = std::max<size_t>((96u << 20) / (planes * sizeof(float)), 1);
    size_t band_rows = std::min(std::max<size_t>(target_px / cols, 1), rows);
    const size_t n_bands = (rows + band_rows - 1) / band_rows;
    const int n_slots = (int)std::min<size_t>(AB_NUM_COPY_STREAMS, n_bands);
    const size_t plane_bytes = ((band_rows * cols * sizeof(float) + 255) / 256) * 256;
    for (int s = 0; s < n_slots; ++s) {
        ab_status st = ab_scratch_reserve(ctx, s, plane_bytes * planes);
        if (st != AB_OK) return st;
    }
    for (size_t b = 0; b < n_bands; ++b) {
        const int slot = (int)(b % n_slots);
        cudaStream_t s = ctx->copy_streams[slot];
        const size_t r0 = b * band_rows, nr = std::min(band_rows, rows - r0);
        const size_t off = r0 * cols, px = nr * cols;
        char *base = static_cast<char *>(ctx->d_scratch[slot]);
        const float *d_frames[AB_MAX_FRAMES];
        size_t k = 0;
        for (size_t f = 0; f < n_frames; ++f, ++k) {
            float *d = reinterpret_cast<float *>(base + k * plane_bytes);
            AB_CUDA(ctx, cudaMemcpyAsync(d, frames[f] + off, px * sizeof(float),
                                         cudaMemcpyHostToDevice, s));
            d_frames[f] = d;
        }
        float *d_bias = nullptr, *d_dark = nullptr;
        if (bias) {
            d_bias = reinterpret_cast<float *>(base + k++ * plane_bytes);
            AB_CUDA(ctx, cudaMemcpyAsync(d_bias, bias + off, px * sizeof(float),
                                         cudaMemcpyHostToDevice, s));
        }
        if (dark) {
            d_dark = reinterpret_cast<float *>(base + k++ * plane_bytes);
            AB_CUDA(ctx, cudaMemcpyAsync(d_dark, dark + off, px * sizeof(float),
                                         cudaMemcpyHostToDevice, s));
        }
        float *d_out = reinterpret_cast<float *>(base + k * plane_bytes);
        ab_status st = ab_median_combine_dev(ctx, d_frames, n_frames, px, d_bias, d_dark, dark_ratio,
                                             d_out, s);
        if (st != AB_OK) return st;
        AB_CUDA(ctx, cudaMemcpyAsync(out + off, d_out, px * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    for (int s = 0; s < n_slots; ++s) AB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_streams[s]));
    return AB_OK;
}

ab_status ab_normalize_flat_dev(ab_ctx *ctx, float *d_data, size_t len, void *stream) {
    if (!ctx || (len && !d_data)) return AB_ERR_INVALID;
    if (len == 0) return AB_OK;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    return normalize_flat_on(ctx, d_data, len, (cudaStream_t)stream);
}

ab_status ab_normalize_flat(ab_ctx *ctx, float *data, size_t len) {
    if (!ctx || (len && !data)) return AB_ERR_INVALID;
    if (len == 0) return AB_OK;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    ab_status st = ab_scratch_reserve(ctx, 0, len * sizeof(float));
    if (st != AB_OK) return st;
    float *d = static_cast<float *>(ctx->d_scratch[0]);
    AB_CUDA(ctx, cudaMemcpyAsync(d, data, len * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    st = normalize_flat_on(ctx, d, len, ctx->stream);
    if (st != AB_OK) return st;
    AB_CUDA(ctx, cudaMemcpyAsync(data, d, len * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}

ab_status ab_calibrate_image_dev(ab_ctx *ctx, const float *d_raw, size_t len,
                                 const ab_calibration_config *cfg, float *d_out, void *stream) {
    if (!ctx || !cfg || (len && (!d_raw || !d_out))) return AB_ERR_INVALID;
    if (len == 0) return AB_OK;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    calibrate_kernel<<<cal_blocks(ctx, len), CAL_THREADS, 0, (cudaStream_t)stream>>>(
        d_raw, len, cfg->master_bias, cfg->master_dark, cfg->master_flat, cfg->dark_exposure_ratio,
        cfg->clamp_negative != 0, d_out);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

ab_status ab_calibrate_image(ab_ctx *ctx, const float *raw, size_t len,
                             const ab_calibration_config *cfg, float *out) {
    if (!ctx || !cfg || (len && (!raw || !out))) return AB_ERR_INVALID;
    if (len == 0) return AB_OK;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    const size_t plane = (len * sizeof(float) + 255) / 256 * 256;
    ab_status st = ab_scratch_reserve(ctx, 0, plane * 5);
    if (st != AB_OK) return st;
    char *base = static_cast<char *>(ctx->d_scratch[0]);
    const float *host[4] = {raw, cfg->master_bias, cfg->master_dark, cfg->master_flat};
    float *dev[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < 4; ++i) {
        if (!host[i]) continue;
        dev[i] = reinterpret_cast<float *>(base + i * plane);
        AB_CUDA(ctx, cudaMemcpyAsync(dev[i], host[i], len * sizeof(float), cudaMemcpyHostToDevice,
                                     ctx->stream));
    }
    float *d_out = reinterpret_cast<float *>(base + 4 * plane);
    calibrate_kernel<<<cal_blocks(ctx, len), CAL_THREADS, 0, ctx->stream>>>(
        dev[0], len, dev[1], dev[2], dev[3], cfg->dark_exposure_ratio, cfg->clamp_negative != 0,
        d_out);
    AB_LAUNCH_CHECK(ctx);
    AB_CUDA(ctx, cudaMemcpyAsync(out, d_out, len * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}

}  // extern "C"
