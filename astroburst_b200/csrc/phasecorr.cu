// phasecorr.cu -- phase-correlation offset estimate: what stack(align=true), the
// GUI default, runs before the sigma-clip stack, and the affine aligner's fallback.
//
// Reference (src-tauri/src/):
//   core/alignment/phase_correlation.rs:22-161  phase_correlate, extract_crop, correlate_single,
//                                               is_constant_or_zero
//   core/alignment/downsample.rs:6-46           area_downsample
//   math/window.rs:3-18 hann_periodic; math/fft.rs:137-167,202-226,247-281 forward/inverse 2-D FFT
//   (inverse scaled by 1/N), prepare_windowed_buffer, extract_real, find_peak;
//   math/complex.rs:17-44 cross_power_spectrum; math/normalization.rs:128-170 compute_mean_sigma,
//   compute_snr; math/subpixel.rs:27-100 quadratic refinement with circular unwrap.
// The reference's FFT is rustfft 6.4.1 (un-vendored third party, f64); here cuFFT Z2Z on the
// same power-of-two grids.  Both are exact DFTs up to f64 rounding, so the correlation surface
// agrees to ~1e-15 and the shifts to ~1e-12 px: parity for this row is tolerance-based.
//
// Per-pixel work on the GPU: the finite min/max scan, the area downsample of both frames (the only
// full-resolution passes, HBM-bound), windowing + zero padding, the cross-power spectrum, scaling.
// The peak search, mean/sigma and the 3-point refinement run on the <= 512 x 512 surface on the
// host in the reference's sequential order.
#include <cufft.h>

#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include <math_constants.h>

#include "ab_internal.h"

namespace {

constexpr size_t COARSE_MAX_DIM = 512;    // phase_correlation.rs:10
constexpr size_t REFINE_CROP_SIZE = 512;  // :11
constexpr double PC_EPSILON = 1e-15;      // :13
constexpr int PC_THREADS = 256;
constexpr int SCAN_BLOCKS = 592;

struct ScanPartial {
    float mn, mx;
    unsigned long long finite;
};

// is_constant_or_zero's scan, phase_correlation.rs:143-160
__global__ void __launch_bounds__(PC_THREADS)
pc_scan_kernel(const float *__restrict__ d, size_t len, ScanPartial *__restrict__ partials) {
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    unsigned long long cnt = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const float v = __ldg(d + i);
        if (fabsf(v) < CUDART_INF_F) {
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
            ++cnt;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_down_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, o));
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    __shared__ ScanPartial sh[PC_THREADS / 32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = {mn, mx, cnt};
    __syncthreads();
    if (threadIdx.x == 0) {
        ScanPartial r = sh[0];
        for (int w = 1; w < PC_THREADS / 32; ++w) {
            r.mn = fminf(r.mn, sh[w].mn);
            r.mx = fmaxf(r.mx, sh[w].mx);
            r.finite += sh[w].finite;
        }
        partials[blockIdx.x] = r;
    }
}

// area_downsample, downsample.rs:6-46: one thread per output pixel, the box summed in the
// reference's order (rows, then columns) in f64
__global__ void __launch_bounds__(PC_THREADS)
area_downsample_kernel(const float *__restrict__ src, size_t in_rows, size_t in_cols,
                       size_t out_rows, size_t out_cols, double scale_y, double scale_x,
                       float *__restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= out_rows * out_cols) return;
    const size_t oy = idx / out_cols, ox = idx % out_cols;
    auto clamp_index = [](long long i, size_t n) -> size_t {  // imaging/boundary.rs:8-20
        return i < 0 ? 0 : ((size_t)i >= n ? n - 1 : (size_t)i);
    };
    const size_t y0 = clamp_index((long long)floor((double)oy * scale_y), in_rows);
    const long long y1r = (long long)ceil((double)(oy + 1) * scale_y);
    const size_t y1 = y1r <= 0 ? 0 : ((size_t)y1r < in_rows ? (size_t)y1r : in_rows);
    const size_t x0 = clamp_index((long long)floor((double)ox * scale_x), in_cols);
    const long long x1r = (long long)ceil((double)(ox + 1) * scale_x);
    const size_t x1 = x1r <= 0 ? 0 : ((size_t)x1r < in_cols ? (size_t)x1r : in_cols);
    double sum = 0.0;
    unsigned count = 0;
    for (size_t y = y0; y < y1; ++y) {
        const float *row = src + y * in_cols;
        for (size_t x = x0; x < x1; ++x) {
            const float v = __ldg(row + x);
            if (fabsf(v) < CUDART_INF_F) {
                sum += (double)v;
                ++count;
            }
        }
    }
    out[idx] = count > 0 ? (float)(sum / (double)count) : 0.0f;
}

// prepare_windowed_buffer, fft.rs:202-226 (img: rows x cols view with row pitch `pitch`)
__global__ void __launch_bounds__(PC_THREADS)
window_kernel(const float *__restrict__ img, size_t rows, size_t cols, size_t pitch,
              const double *__restrict__ win_y, const double *__restrict__ win_x, size_t fft_rows,
              size_t fft_cols, cufftDoubleComplex *__restrict__ buf) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= fft_rows * fft_cols) return;
    const size_t y = idx / fft_cols, x = idx % fft_cols;
    double w = 0.0;
    if (y < rows && x < cols) {
        const double v = (double)img[y * pitch + x];
        if (fabs(v) < CUDART_INF) w = v * win_y[y] * win_x[x];
    }
    buf[idx] = make_cuDoubleComplex(w, 0.0);
}

// cross_power_spectrum, complex.rs:17-44
__global__ void __launch_bounds__(PC_THREADS)
cross_power_kernel(const cufftDoubleComplex *__restrict__ fa, const cufftDoubleComplex *__restrict__ fb,
                   size_t n, double epsilon, cufftDoubleComplex *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const cufftDoubleComplex a = fa[i], b = fb[i];
    const double re = a.x * b.x + a.y * b.y;
    const double im = a.y * b.x - a.x * b.y;
    const double mag = sqrt(re * re + im * im);
    out[i] = mag > epsilon ? make_cuDoubleComplex(re / mag, im / mag) : make_cuDoubleComplex(0.0, 0.0);
}

// inverse_2d's 1/N scaling (fft.rs:162-166) + extract_real (fft.rs:247-249)
__global__ void __launch_bounds__(PC_THREADS)
scale_real_kernel(const cufftDoubleComplex *__restrict__ c, size_t n, double norm,
                  double *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = c[i].x * norm;
}

// ---- per-context resources (plans, scratch); created lazily, released by ab_ctx_destroy
struct PcState {
    std::map<unsigned long long, cufftHandle> plans;  // (rows << 32 | cols) -> Z2Z plan
    cufftDoubleComplex *fa = nullptr, *fb = nullptr, *cross = nullptr;
    double *surface = nullptr, *win = nullptr;  // win: [2][COARSE_MAX_DIM]
    float *ds[2] = {nullptr, nullptr};
    ScanPartial *partials = nullptr;
    double *h_surface = nullptr;  // pinned
    ScanPartial *h_partials = nullptr;
};
std::mutex g_mu;
std::map<ab_ctx *, PcState *> g_states;

constexpr size_t MAX_FFT = 512 * 512;  // correlate_single never sees more than 512 x 512

ab_status get_state(ab_ctx *ctx, PcState **out) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_states.find(ctx);
    if (it != g_states.end()) {
        *out = it->second;
        return AB_OK;
    }
    PcState *st = new PcState();
    bool ok = cudaMalloc(&st->fa, MAX_FFT * sizeof(cufftDoubleComplex)) == cudaSuccess;
    ok = ok && cudaMalloc(&st->fb, MAX_FFT * sizeof(cufftDoubleComplex)) == cudaSuccess;
    ok = ok && cudaMalloc(&st->cross, MAX_FFT * sizeof(cufftDoubleComplex)) == cudaSuccess;
    ok = ok && cudaMalloc(&st->surface, MAX_FFT * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&st->win, 2 * COARSE_MAX_DIM * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&st->ds[0], MAX_FFT * sizeof(float)) == cudaSuccess;
    ok = ok && cudaMalloc(&st->ds[1], MAX_FFT * sizeof(float)) == cudaSuccess;
    ok = ok && cudaMalloc(&st->partials, 2 * SCAN_BLOCKS * sizeof(ScanPartial)) == cudaSuccess;
    ok = ok && cudaMallocHost(&st->h_surface, MAX_FFT * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMallocHost(&st->h_partials, 2 * SCAN_BLOCKS * sizeof(ScanPartial)) == cudaSuccess;
    if (!ok) {
        delete st;
        return ab_fail(ctx, AB_ERR_NOMEM, "phase correlation: scratch allocation failed");
    }
    g_states[ctx] = st;
    *out = st;
    return AB_OK;
}

ab_status get_plan(ab_ctx *ctx, PcState *st, size_t rows, size_t cols, cufftHandle *plan) {
    const unsigned long long key = ((unsigned long long)rows << 32) | cols;
    auto it = st->plans.find(key);
    if (it != st->plans.end()) {
        *plan = it->second;
        return AB_OK;
    }
    cufftHandle h;
    if (cufftPlan2d(&h, (int)rows, (int)cols, CUFFT_Z2Z) != CUFFT_SUCCESS)
        return ab_fail(ctx, AB_ERR_CUDA, "cufftPlan2d(%zu, %zu) failed", rows, cols);
    st->plans[key] = h;
    *plan = h;
    return AB_OK;
}

size_t next_pow2(size_t n) {
    size_t p = 1;
    while (p < n) p <<= 1;
    return p;
}

unsigned blocks_for(size_t n) { return (unsigned)((n + PC_THREADS - 1) / PC_THREADS); }

// correlate_single, phase_correlation.rs:114-141 on two rows x cols views (rows, cols <= 512)
ab_status correlate_single(ab_ctx *ctx, PcState *st, const float *a, size_t pitch_a, const float *b,
                           size_t pitch_b, size_t rows, size_t cols, cudaStream_t s,
                           ab_phase_result *out) {
    const size_t fr = next_pow2(rows), fc = next_pow2(cols);
    if (fr * fc > MAX_FFT)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "correlate_single: %zu x %zu", rows, cols);
    // hann_periodic, window.rs:3-18 (host libm cos, like the reference; uploaded)
    std::vector<double> win(2 * COARSE_MAX_DIM, 0.0);
    auto hann = [](double *w, size_t n) {
        if (n == 1) {
            w[0] = 1.0;
            return;
        }
        const double two_pi = 2.0 * 3.14159265358979323846, nf = (double)n;
        for (size_t i = 0; i < n; ++i) w[i] = 0.5 * (1.0 - std::cos(two_pi * (double)i / nf));
    };
    hann(win.data(), rows);
    hann(win.data() + COARSE_MAX_DIM, cols);
    AB_CUDA(ctx, cudaMemcpyAsync(st->win, win.data(), win.size() * sizeof(double),
                                 cudaMemcpyHostToDevice, s));
    AB_CUDA(ctx, cudaStreamSynchronize(s));  // `win` is a stack-lifetime host buffer
    const size_t n = fr * fc;
    window_kernel<<<blocks_for(n), PC_THREADS, 0, s>>>(a, rows, cols, pitch_a, st->win,
                                                       st->win + COARSE_MAX_DIM, fr, fc, st->fa);
    AB_LAUNCH_CHECK(ctx);
    window_kernel<<<blocks_for(n), PC_THREADS, 0, s>>>(b, rows, cols, pitch_b, st->win,
                                                       st->win + COARSE_MAX_DIM, fr, fc, st->fb);
    AB_LAUNCH_CHECK(ctx);
    cufftHandle plan;
    ab_status rc = get_plan(ctx, st, fr, fc, &plan);
    if (rc != AB_OK) return rc;
    if (cufftSetStream(plan, s) != CUFFT_SUCCESS ||
        cufftExecZ2Z(plan, st->fa, st->fa, CUFFT_FORWARD) != CUFFT_SUCCESS ||
        cufftExecZ2Z(plan, st->fb, st->fb, CUFFT_FORWARD) != CUFFT_SUCCESS)
        return ab_fail(ctx, AB_ERR_CUDA, "cufftExecZ2Z (forward) failed");
    cross_power_kernel<<<blocks_for(n), PC_THREADS, 0, s>>>(st->fa, st->fb, n, PC_EPSILON, st->cross);
    AB_LAUNCH_CHECK(ctx);
    if (cufftExecZ2Z(plan, st->cross, st->cross, CUFFT_INVERSE) != CUFFT_SUCCESS)
        return ab_fail(ctx, AB_ERR_CUDA, "cufftExecZ2Z (inverse) failed");
    scale_real_kernel<<<blocks_for(n), PC_THREADS, 0, s>>>(st->cross, n, 1.0 / (double)n, st->surface);
    AB_LAUNCH_CHECK(ctx);
    AB_CUDA(ctx, cudaMemcpyAsync(st->h_surface, st->surface, n * sizeof(double),
                                 cudaMemcpyDeviceToHost, s));
    AB_CUDA(ctx, cudaStreamSynchronize(s));
    const double *sf = st->h_surface;
    // find_peak, fft.rs:271-281 (first maximum)
    size_t best = 0;
    for (size_t i = 1; i < n; ++i)
        if (sf[i] > sf[best]) best = i;
    const size_t peak_y = best / fc, peak_x = best % fc;
    const double peak_val = sf[best];
    // compute_mean_sigma, normalization.rs:128-163; compute_snr :165-170
    double sum = 0.0, count = 0.0;
    for (size_t i = 0; i < n; ++i)
        if (std::isfinite(sf[i])) {
            sum += sf[i];
            count += 1.0;
        }
    double mean = 0.0, sigma = 0.0;
    if (count >= 1.0) {
        mean = sum / count;
        double var_sum = 0.0;
        for (size_t i = 0; i < n; ++i)
            if (std::isfinite(sf[i])) {
                const double d = sf[i] - mean;
                var_sum += d * d;
            }
        sigma = std::sqrt(var_sum / (count > 1.0 ? count - 1.0 : 1.0));
    }
    const double eps = 1e-15;  // FftFloat::epsilon_val() for f64 (math/fft.rs:57), NOT f64::EPSILON
    const double confidence = std::fabs(sigma) < eps ? 0.0 : (peak_val - mean) / sigma;
    // unwrap_and_refine, subpixel.rs:27-100
    auto unwrap = [](size_t peak, size_t size) {
        return peak > size / 2 ? (double)peak - (double)size : (double)peak;
    };
    auto refine = [&](double prev, double center, double next) {
        const double denom = 2.0 * (2.0 * center - prev - next);
        if (std::fabs(denom) < eps) return 0.0;
        const double r = (prev - next) / denom;
        return std::fmin(std::fmax(r, -0.5), 0.5);
    };
    const size_t py = peak_y == 0 ? fr - 1 : peak_y - 1, ny = peak_y == fr - 1 ? 0 : peak_y + 1;
    const size_t px = peak_x == 0 ? fc - 1 : peak_x - 1, nx = peak_x == fc - 1 ? 0 : peak_x + 1;
    const double c = sf[peak_y * fc + peak_x];
    out->dy = unwrap(peak_y, fr) + refine(sf[py * fc + peak_x], c, sf[ny * fc + peak_x]);
    out->dx = unwrap(peak_x, fc) + refine(sf[peak_y * fc + px], c, sf[peak_y * fc + nx]);
    out->confidence = confidence;
    return AB_OK;
}

ab_status downsample_on(ab_ctx *ctx, const float *d_src, size_t in_rows, size_t in_cols,
                        size_t out_rows, size_t out_cols, float *d_out, cudaStream_t s) {
    if (in_rows == out_rows && in_cols == out_cols) {  // downsample.rs:9-11
        AB_CUDA(ctx, cudaMemcpyAsync(d_out, d_src, in_rows * in_cols * sizeof(float),
                                     cudaMemcpyDeviceToDevice, s));
        return AB_OK;
    }
    const double sy = (double)in_rows / (double)out_rows, sx = (double)in_cols / (double)out_cols;
    area_downsample_kernel<<<blocks_for(out_rows * out_cols), PC_THREADS, 0, s>>>(
        d_src, in_rows, in_cols, out_rows, out_cols, sy, sx, d_out);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

}  // namespace

void ab_phasecorr_release(ab_ctx *ctx) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_states.find(ctx);
    if (it == g_states.end()) return;
    PcState *st = it->second;
    for (auto &p : st->plans) cufftDestroy(p.second);
    cudaFree(st->fa);
    cudaFree(st->fb);
    cudaFree(st->cross);
    cudaFree(st->surface);
    cudaFree(st->win);
    cudaFree(st->ds[0]);
    cudaFree(st->ds[1]);
    cudaFree(st->partials);
    cudaFreeHost(st->h_surface);
    cudaFreeHost(st->h_partials);
    delete st;
    g_states.erase(it);
}

// phase_correlate on two device planes of the same size (rows x cols, contiguous).  Synchronizes
// `stream` (the result is a host struct; the surface post-processing is host scalar code).
ab_status ab_phase_correlate_on(ab_ctx *ctx, const float *d_ref, const float *d_tgt, size_t rows,
                                size_t cols, cudaStream_t s, ab_phase_result *out) {
    out->dx = out->dy = out->confidence = 0.0;
    if (rows == 0 || cols == 0) return AB_OK;
    PcState *st;
    ab_status rc = get_state(ctx, &st);
    if (rc != AB_OK) return rc;
    // is_constant_or_zero, phase_correlation.rs:42-48,143-160
    const size_t len = rows * cols;
    const int sb = (int)std::min<size_t>(SCAN_BLOCKS, (len + PC_THREADS - 1) / PC_THREADS);
    pc_scan_kernel<<<sb, PC_THREADS, 0, s>>>(d_ref, len, st->partials);
    AB_LAUNCH_CHECK(ctx);
    pc_scan_kernel<<<sb, PC_THREADS, 0, s>>>(d_tgt, len, st->partials + SCAN_BLOCKS);
    AB_LAUNCH_CHECK(ctx);
    AB_CUDA(ctx, cudaMemcpyAsync(st->h_partials, st->partials, 2 * SCAN_BLOCKS * sizeof(ScanPartial),
                                 cudaMemcpyDeviceToHost, s));
    AB_CUDA(ctx, cudaStreamSynchronize(s));
    for (int img = 0; img < 2; ++img) {
        float mn = INFINITY, mx = -INFINITY;
        unsigned long long fin = 0;
        for (int i = 0; i < sb; ++i) {
            const ScanPartial &p = st->h_partials[img * SCAN_BLOCKS + i];
            mn = std::fmin(mn, p.mn);
            mx = std::fmax(mx, p.mx);
            fin += p.finite;
        }
        if (fin < 16 || std::fabs(mx - mn) < 1e-10f) return AB_OK;  // zeros
    }
    if (rows <= COARSE_MAX_DIM && cols <= COARSE_MAX_DIM)
        return correlate_single(ctx, st, d_ref, cols, d_tgt, cols, rows, cols, s, out);

    const double scale_y = (double)rows / (double)COARSE_MAX_DIM;
    const double scale_x = (double)cols / (double)COARSE_MAX_DIM;
    const size_t ds_rows = std::min(COARSE_MAX_DIM, rows), ds_cols = std::min(COARSE_MAX_DIM, cols);
    if ((rc = downsample_on(ctx, d_ref, rows, cols, ds_rows, ds_cols, st->ds[0], s)) != AB_OK) return rc;
    if ((rc = downsample_on(ctx, d_tgt, rows, cols, ds_rows, ds_cols, st->ds[1], s)) != AB_OK) return rc;
    ab_phase_result coarse;
    rc = correlate_single(ctx, st, st->ds[0], ds_cols, st->ds[1], ds_cols, ds_rows, ds_cols, s, &coarse);
    if (rc != AB_OK) return rc;
    const double coarse_dx = coarse.dx * scale_x, coarse_dy = coarse.dy * scale_y;

    const size_t half = REFINE_CROP_SIZE / 2;
    const size_t ref_cy = rows / 2, ref_cx = cols / 2;
    auto clampi = [](double v, size_t n) -> size_t {  // `.round() as isize).clamp(0, n - 1)`
        const double r = std::round(v);
        if (!(r > 0.0)) return 0;
        return r >= (double)(n - 1) ? n - 1 : (size_t)r;
    };
    const size_t tgt_cy = clampi((double)ref_cy + coarse_dy, rows);
    const size_t tgt_cx = clampi((double)ref_cx + coarse_dx, cols);
    struct Crop {
        size_t y0, y1, x0, x1;
    };
    auto crop = [&](size_t cy, size_t cx) {  // extract_crop, :101-112
        return Crop{cy > half ? cy - half : 0, std::min(cy + half, rows), cx > half ? cx - half : 0,
                    std::min(cx + half, cols)};
    };
    const Crop cr = crop(ref_cy, ref_cx), ct = crop(tgt_cy, tgt_cx);
    if (cr.y1 - cr.y0 != ct.y1 - ct.y0 || cr.x1 - cr.x0 != ct.x1 - ct.x0) {  // :82-88
        out->dx = coarse_dx;
        out->dy = coarse_dy;
        out->confidence = coarse.confidence;
        return AB_OK;
    }
    ab_phase_result refine;
    rc = correlate_single(ctx, st, d_ref + cr.y0 * cols + cr.x0, cols, d_tgt + ct.y0 * cols + ct.x0,
                          cols, cr.y1 - cr.y0, cr.x1 - cr.x0, s, &refine);
    if (rc != AB_OK) return rc;
    out->dx = coarse_dx + refine.dx;
    out->dy = coarse_dy + refine.dy;
    out->confidence = refine.confidence;
    return AB_OK;
}

extern "C" {

ab_status ab_phase_correlate_dev(ab_ctx *ctx, const float *d_ref, const float *d_tgt, size_t rows,
                                 size_t cols, ab_phase_result *out, void *stream) {
    if (!ctx || !out || ((rows * cols) && (!d_ref || !d_tgt))) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    return ab_phase_correlate_on(ctx, d_ref, d_tgt, rows, cols, (cudaStream_t)stream, out);
}

ab_status ab_phase_correlate(ab_ctx *ctx, const float *ref, size_t ref_rows, size_t ref_cols,
                             const float *tgt, size_t tgt_rows, size_t tgt_cols,
                             ab_phase_result *out) {
    if (!ctx || !out || !ref || !tgt) return AB_ERR_INVALID;
    const size_t rows = std::min(ref_rows, tgt_rows), cols = std::min(ref_cols, tgt_cols);  // :26-41
    out->dx = out->dy = out->confidence = 0.0;
    if (rows == 0 || cols == 0) return AB_OK;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    const size_t plane = (rows * cols * sizeof(float) + 255) / 256 * 256;
    ab_status st = ab_scratch_reserve(ctx, 0, plane * 2);
    if (st != AB_OK) return st;
    float *d_ref = static_cast<float *>(ctx->d_scratch[0]);
    float *d_tgt = reinterpret_cast<float *>(static_cast<char *>(ctx->d_scratch[0]) + plane);
    AB_CUDA(ctx, cudaMemcpy2DAsync(d_ref, cols * sizeof(float), ref, ref_cols * sizeof(float),
                                   cols * sizeof(float), rows, cudaMemcpyHostToDevice, ctx->stream));
    AB_CUDA(ctx, cudaMemcpy2DAsync(d_tgt, cols * sizeof(float), tgt, tgt_cols * sizeof(float),
                                   cols * sizeof(float), rows, cudaMemcpyHostToDevice, ctx->stream));
    return ab_phase_correlate_on(ctx, d_ref, d_tgt, rows, cols, ctx->stream, out);
}

ab_status ab_area_downsample_dev(ab_ctx *ctx, const float *d_src, size_t in_rows, size_t in_cols,
                                 size_t out_rows, size_t out_cols, float *d_out, void *stream) {
    if (!ctx || !d_src || !d_out || !in_rows || !in_cols || !out_rows || !out_cols)
        return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    return downsample_on(ctx, d_src, in_rows, in_cols, out_rows, out_cols, d_out,
                         (cudaStream_t)stream);
}

ab_status ab_area_downsample(ab_ctx *ctx, const float *src, size_t in_rows, size_t in_cols,
                             size_t out_rows, size_t out_cols, float *out) {
    if (!ctx || !src || !out || !in_rows || !in_cols || !out_rows || !out_cols) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    const size_t in_b = (in_rows * in_cols * sizeof(float) + 255) / 256 * 256;
    ab_status st = ab_scratch_reserve(ctx, 0, in_b + out_rows * out_cols * sizeof(float));
    if (st != AB_OK) return st;
    float *d_in = static_cast<float *>(ctx->d_scratch[0]);
    float *d_out = reinterpret_cast<float *>(static_cast<char *>(ctx->d_scratch[0]) + in_b);
    AB_CUDA(ctx, cudaMemcpyAsync(d_in, src, in_rows * in_cols * sizeof(float), cudaMemcpyHostToDevice,
                                 ctx->stream));
    st = downsample_on(ctx, d_in, in_rows, in_cols, out_rows, out_cols, d_out, ctx->stream);
    if (st != AB_OK) return st;
    AB_CUDA(ctx, cudaMemcpyAsync(out, d_out, out_rows * out_cols * sizeof(float),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}

}  // extern "C"
