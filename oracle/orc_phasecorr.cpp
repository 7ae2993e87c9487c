// orc_phasecorr.cpp -- CPU oracle (TEST INFRASTRUCTURE, see oracle.h) for the phase-correlation
// offset estimate.
//
// Restates (paths relative to src-tauri/src/):
//   core/alignment/phase_correlation.rs:22-161  phase_correlate, extract_crop, correlate_single,
//                                               is_constant_or_zero
//   core/alignment/downsample.rs:6-46           area_downsample
//   math/window.rs:3-18        hann_periodic
//   math/fft.rs:137-167        forward_2d / inverse_2d (rows, then columns; inverse scaled by 1/N)
//   math/fft.rs:202-226,247-281 prepare_windowed_buffer, extract_real, find_peak
//   math/complex.rs:17-44      cross_power_spectrum
//   math/normalization.rs:128-170 compute_mean_sigma, compute_snr
//   math/subpixel.rs:27-100    quadratic_refine_1d, unwrap_circular_peak, unwrap_and_refine
// Third-party arithmetic: the reference's 1-D FFTs are rustfft 6.4.1 (f64), whose source is not
// under /root/reference.  Its published algorithm for power-of-two lengths is a radix-4/8
// Cooley-Tukey DFT; restated here as an iterative radix-2 DFT.  Every exact DFT agrees to f64
// rounding, so this oracle pins the result to ~1e-12 px, not bit-exactly: PARITY UNPINNED at the
// bit level for this row; anchored on the reference's own tolerance tests
// (phase_correlation.rs:197-240, +-0.5..1 px).
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

#include "oracle.h"

namespace {
typedef std::complex<double> cd;
const double PI = 3.14159265358979323846;

void fft1d(cd *a, size_t n, bool inverse) {  // n a power of two
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = 2.0 * PI / (double)len * (inverse ? 1.0 : -1.0);
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const cd w(std::cos(ang * (double)k), std::sin(ang * (double)k));
                const cd u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}
void fft2d(std::vector<cd> &buf, size_t rows, size_t cols, bool inverse) {
#pragma omp parallel for
    for (long long y = 0; y < (long long)rows; ++y) fft1d(&buf[(size_t)y * cols], cols, inverse);
#pragma omp parallel for
    for (long long x = 0; x < (long long)cols; ++x) {
        std::vector<cd> col(rows);
        for (size_t y = 0; y < rows; ++y) col[y] = buf[y * cols + (size_t)x];
        fft1d(col.data(), rows, inverse);
        for (size_t y = 0; y < rows; ++y) buf[y * cols + (size_t)x] = col[y];
    }
    if (inverse) {
        const double norm = 1.0 / (double)(rows * cols);
        for (cd &c : buf) c = cd(c.real() * norm, c.imag() * norm);
    }
}
size_t next_pow2(size_t n) {
    size_t p = 1;
    while (p < n) p <<= 1;
    return p;
}
std::vector<double> hann_periodic(size_t n) {
    if (n == 0) return {};
    if (n == 1) return {1.0};
    std::vector<double> w(n);
    const double two_pi = 2.0 * PI, nf = (double)n;
    for (size_t i = 0; i < n; ++i) w[i] = 0.5 * (1.0 - std::cos(two_pi * (double)i / nf));
    return w;
}
struct View {  // rows x cols window into a row-major image
    const float *p;
    size_t rows, cols, pitch;
    float at(size_t y, size_t x) const { return p[y * pitch + x]; }
};
std::vector<cd> windowed(const View &v, const std::vector<double> &wy, const std::vector<double> &wx,
                         size_t fr, size_t fc) {
    std::vector<cd> buf(fr * fc, cd(0.0, 0.0));
    for (size_t y = 0; y < v.rows; ++y)
        for (size_t x = 0; x < v.cols; ++x) {
            const double val = (double)v.at(y, x);
            buf[y * fc + x] = cd(std::isfinite(val) ? val * wy[y] * wx[x] : 0.0, 0.0);
        }
    return buf;
}
bool is_constant_or_zero(const View &v) {
    float mn = INFINITY, mx = -INFINITY;
    unsigned long long fin = 0;
    for (size_t y = 0; y < v.rows; ++y)
        for (size_t x = 0; x < v.cols; ++x) {
            const float p = v.at(y, x);
            if (std::isfinite(p)) {
                if (p < mn) mn = p;
                if (p > mx) mx = p;
                ++fin;
            }
        }
    return fin < 16 || std::fabs(mx - mn) < 1e-10f;
}
void correlate_single(const View &a, const View &b, double *dx, double *dy, double *conf) {
    const size_t rows = a.rows, cols = a.cols, fr = next_pow2(rows), fc = next_pow2(cols);
    const std::vector<double> wy = hann_periodic(rows), wx = hann_periodic(cols);
    std::vector<cd> fa = windowed(a, wy, wx, fr, fc), fb = windowed(b, wy, wx, fr, fc);
    fft2d(fa, fr, fc, false);
    fft2d(fb, fr, fc, false);
    std::vector<cd> cross(fr * fc);
    for (size_t i = 0; i < cross.size(); ++i) {
        const double re = fa[i].real() * fb[i].real() + fa[i].imag() * fb[i].imag();
        const double im = fa[i].imag() * fb[i].real() - fa[i].real() * fb[i].imag();
        const double mag = std::sqrt(re * re + im * im);
        cross[i] = mag > 1e-15 ? cd(re / mag, im / mag) : cd(0.0, 0.0);
    }
    fft2d(cross, fr, fc, true);
    const size_t n = fr * fc;
    std::vector<double> sf(n);
    for (size_t i = 0; i < n; ++i) sf[i] = cross[i].real();
    size_t best = 0;
    for (size_t i = 1; i < n; ++i)
        if (sf[i] > sf[best]) best = i;
    const size_t peak_y = best / fc, peak_x = best % fc;
    double sum = 0.0, count = 0.0;
    for (double v : sf)
        if (std::isfinite(v)) {
            sum += v;
            count += 1.0;
        }
    double mean = 0.0, sigma = 0.0;
    if (count >= 1.0) {
        mean = sum / count;
        double var_sum = 0.0;
        for (double v : sf)
            if (std::isfinite(v)) var_sum += (v - mean) * (v - mean);
        sigma = std::sqrt(var_sum / (count > 1.0 ? count - 1.0 : 1.0));
    }
    const double eps = 1e-15;  // FftFloat::epsilon_val() for f64, math/fft.rs:57 (normalization.rs:166, subpixel.rs:55)
    *conf = std::fabs(sigma) < eps ? 0.0 : (sf[best] - mean) / sigma;
    auto unwrap = [](size_t peak, size_t size) {
        return peak > size / 2 ? (double)peak - (double)size : (double)peak;
    };
    auto refine = [&](double prev, double center, double next) {
        const double denom = 2.0 * (2.0 * center - prev - next);
        if (std::fabs(denom) < eps) return 0.0;
        return std::fmin(std::fmax((prev - next) / denom, -0.5), 0.5);
    };
    const size_t py = peak_y == 0 ? fr - 1 : peak_y - 1, ny = peak_y == fr - 1 ? 0 : peak_y + 1;
    const size_t px = peak_x == 0 ? fc - 1 : peak_x - 1, nx = peak_x == fc - 1 ? 0 : peak_x + 1;
    const double c = sf[peak_y * fc + peak_x];
    *dy = unwrap(peak_y, fr) + refine(sf[py * fc + peak_x], c, sf[ny * fc + peak_x]);
    *dx = unwrap(peak_x, fc) + refine(sf[peak_y * fc + px], c, sf[peak_y * fc + nx]);
}
}  // namespace

extern "C" {

void orc_area_downsample(const float *src, size_t in_rows, size_t in_cols, size_t out_rows,
                         size_t out_cols, float *out) {
    if (in_rows == out_rows && in_cols == out_cols) {
        std::copy(src, src + in_rows * in_cols, out);
        return;
    }
    const double scale_y = (double)in_rows / (double)out_rows, scale_x = (double)in_cols / (double)out_cols;
    auto clamp_index = [](long long i, size_t n) -> size_t {
        return i < 0 ? 0 : ((size_t)i >= n ? n - 1 : (size_t)i);
    };
#pragma omp parallel for
    for (long long oy = 0; oy < (long long)out_rows; ++oy) {
        const size_t y0 = clamp_index((long long)std::floor((double)oy * scale_y), in_rows);
        const long long y1r = (long long)std::ceil((double)(oy + 1) * scale_y);
        const size_t y1 = y1r <= 0 ? 0 : std::min((size_t)y1r, in_rows);
        for (size_t ox = 0; ox < out_cols; ++ox) {
            const size_t x0 = clamp_index((long long)std::floor((double)ox * scale_x), in_cols);
            const long long x1r = (long long)std::ceil((double)(ox + 1) * scale_x);
            const size_t x1 = x1r <= 0 ? 0 : std::min((size_t)x1r, in_cols);
            double sum = 0.0;
            unsigned count = 0;
            for (size_t y = y0; y < y1; ++y)
                for (size_t x = x0; x < x1; ++x) {
                    const float v = src[y * in_cols + x];
                    if (std::isfinite(v)) {
                        sum += (double)v;
                        ++count;
                    }
                }
            out[(size_t)oy * out_cols + ox] = count > 0 ? (float)(sum / (double)count) : 0.0f;
        }
    }
}

void orc_phase_correlate(const float *ref, size_t ref_rows, size_t ref_cols, const float *tgt,
                         size_t tgt_rows, size_t tgt_cols, double *dx, double *dy, double *conf) {
    *dx = *dy = *conf = 0.0;
    const size_t rows = std::min(ref_rows, tgt_rows), cols = std::min(ref_cols, tgt_cols);
    if (rows == 0 || cols == 0) return;
    const View r{ref, rows, cols, ref_cols}, t{tgt, rows, cols, tgt_cols};
    if (is_constant_or_zero(r) || is_constant_or_zero(t)) return;
    const size_t CM = 512, RC = 512;
    if (rows <= CM && cols <= CM) {
        correlate_single(r, t, dx, dy, conf);
        return;
    }
    // contiguous cropped copies for the downsample
    std::vector<float> rc(rows * cols), tc(rows * cols);
    for (size_t y = 0; y < rows; ++y) {
        std::copy(ref + y * ref_cols, ref + y * ref_cols + cols, rc.begin() + y * cols);
        std::copy(tgt + y * tgt_cols, tgt + y * tgt_cols + cols, tc.begin() + y * cols);
    }
    const double scale_y = (double)rows / (double)CM, scale_x = (double)cols / (double)CM;
    const size_t ds_rows = std::min(CM, rows), ds_cols = std::min(CM, cols);
    std::vector<float> rd(ds_rows * ds_cols), td(ds_rows * ds_cols);
    orc_area_downsample(rc.data(), rows, cols, ds_rows, ds_cols, rd.data());
    orc_area_downsample(tc.data(), rows, cols, ds_rows, ds_cols, td.data());
    double cdx, cdy, cconf;
    correlate_single(View{rd.data(), ds_rows, ds_cols, ds_cols}, View{td.data(), ds_rows, ds_cols, ds_cols},
                     &cdx, &cdy, &cconf);
    const double coarse_dx = cdx * scale_x, coarse_dy = cdy * scale_y;
    const size_t half = RC / 2, ref_cy = rows / 2, ref_cx = cols / 2;
    auto clampi = [](double v, size_t n) -> size_t {
        const double r = std::round(v);
        if (!(r > 0.0)) return 0;
        return r >= (double)(n - 1) ? n - 1 : (size_t)r;
    };
    const size_t tgt_cy = clampi((double)ref_cy + coarse_dy, rows), tgt_cx = clampi((double)ref_cx + coarse_dx, cols);
    auto crop = [&](const std::vector<float> &img, size_t cy, size_t cx) {
        const size_t y0 = cy > half ? cy - half : 0, y1 = std::min(cy + half, rows);
        const size_t x0 = cx > half ? cx - half : 0, x1 = std::min(cx + half, cols);
        return View{img.data() + y0 * cols + x0, y1 - y0, x1 - x0, cols};
    };
    const View cr = crop(rc, ref_cy, ref_cx), ct = crop(tc, tgt_cy, tgt_cx);
    if (cr.rows != ct.rows || cr.cols != ct.cols) {
        *dx = coarse_dx;
        *dy = coarse_dy;
        *conf = cconf;
        return;
    }
    double rdx, rdy, rconf;
    correlate_single(cr, ct, &rdx, &rdy, &rconf);
    *dx = coarse_dx + rdx;
    *dy = coarse_dy + rdy;
    *conf = rconf;
}

}  // extern "C"
