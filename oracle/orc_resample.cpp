// orc_resample.cpp -- CPU oracle (TEST INFRASTRUCTURE, see oracle.h) for the
// Catmull-Rom bicubic sampler and its three callers.
//
// Restates (paths relative to src-tauri/src/):
//   core/imaging/boundary.rs:8-20      clamp_index
//   core/imaging/sampling.rs:4-13      catmull_rom (Keys a = -0.5)
//   core/imaging/sampling.rs:26-48     bilinear_sample (only used by the reference's tests)
//   core/imaging/sampling.rs:51-80     bicubic_sample
//   core/alignment/affine.rs:663-690   warp_image   (0 <= s < dim-1 else 0)
//   core/stacking/align.rs:36-57       shift_image_subpixel (-0.5 <= s <= dim-0.5 else 0)
//   core/imaging/resample.rs:25-61     resample_image (never skips, clamped taps)
#include <cmath>
#include <cstring>

#include "oracle.h"

namespace {

inline size_t clamp_index(long long idx, size_t len) {
    if (len == 0) return 0;
    if (idx < 0) return 0;
    if (idx >= (long long)len) return len - 1;
    return (size_t)idx;
}

// Rust `f64 as i64` saturates; coordinates here are always far inside i64 range
// except NaN -> 0.
inline long long f2i64(double x) {
    if (std::isnan(x)) return 0;
    if (x >= 9.2233720368547758e18) return INT64_MAX;
    if (x <= -9.2233720368547758e18) return INT64_MIN;
    return (long long)x;
}

}  // namespace

extern "C" {

double orc_catmull_rom(double t) {
    const double a = std::fabs(t);
    if (a <= 1.0) return a * a * (1.5 * a - 2.5) + 1.0;
    if (a <= 2.0) return a * (a * (2.5 - 0.5 * a) - 4.0) + 2.0;
    return 0.0;
}

float orc_bilinear_sample(const float *s, size_t rows, size_t cols, double y, double x) {
    if (rows == 0 || cols == 0) return 0.0f;
    const long long ix0 = f2i64(std::floor(x));
    const long long iy0 = f2i64(std::floor(y));
    const double fx = x - (double)ix0;
    const double fy = y - (double)iy0;
    const size_t r0 = clamp_index(iy0, rows), r1 = clamp_index(iy0 + 1, rows);
    const size_t c0 = clamp_index(ix0, cols), c1 = clamp_index(ix0 + 1, cols);
    const double v00 = s[r0 * cols + c0], v01 = s[r0 * cols + c1];
    const double v10 = s[r1 * cols + c0], v11 = s[r1 * cols + c1];
    const double top = v00 + (v01 - v00) * fx;
    const double bot = v10 + (v11 - v10) * fx;
    return (float)(top + (bot - top) * fy);
}

float orc_bicubic_sample(const float *s, size_t rows, size_t cols, double y, double x) {
    if (rows == 0 || cols == 0) return 0.0f;
    const long long ix = f2i64(std::floor(x));
    const long long iy = f2i64(std::floor(y));
    const double fx = x - (double)ix;
    const double fy = y - (double)iy;
    const double wx[4] = {orc_catmull_rom(fx + 1.0), orc_catmull_rom(fx), orc_catmull_rom(fx - 1.0),
                          orc_catmull_rom(fx - 2.0)};
    double val = 0.0;
    for (long long j = 0; j < 4; ++j) {
        const size_t r = clamp_index(iy + j - 1, rows);
        const size_t row_off = r * cols;
        double row_val = 0.0;
        for (long long i = 0; i < 4; ++i) {
            const size_t c = clamp_index(ix + i - 1, cols);
            row_val += (double)s[row_off + c] * wx[i];
        }
        val += row_val * orc_catmull_rom(fy - (double)(j - 1));
    }
    return (float)val;
}

void orc_warp_image(const float *img, size_t src_rows, size_t src_cols, const orc_affine *t,
                    size_t out_rows, size_t out_cols, float *out) {
    std::memset(out, 0, out_rows * out_cols * sizeof(float));
#pragma omp parallel for schedule(dynamic, 8)
    for (long long y = 0; y < (long long)out_rows; ++y) {
        float *row = out + (size_t)y * out_cols;
        for (size_t x = 0; x < out_cols; ++x) {
            const double sx = t->a * (double)x + t->b * (double)y + t->tx;
            const double sy = t->c * (double)x + t->d * (double)y + t->ty;
            // (src_cols - 1) is usize arithmetic in the reference; dims >= 1 here.
            if (sx >= 0.0 && sy >= 0.0 && sx < (double)(src_cols - 1) &&
                sy < (double)(src_rows - 1))
                row[x] = orc_bicubic_sample(img, src_rows, src_cols, sy, sx);
        }
    }
}

void orc_shift_image_subpixel(const float *img, size_t rows, size_t cols, double dy, double dx,
                              float *out) {
    if (std::fabs(dy) < 1e-12 && std::fabs(dx) < 1e-12) {
        std::memcpy(out, img, rows * cols * sizeof(float));
        return;
    }
    std::memset(out, 0, rows * cols * sizeof(float));
    const double rows_f = (double)rows, cols_f = (double)cols;
#pragma omp parallel for schedule(dynamic, 8)
    for (long long y = 0; y < (long long)rows; ++y) {
        float *row = out + (size_t)y * cols;
        for (size_t x = 0; x < cols; ++x) {
            const double sy = (double)y + dy;
            const double sx = (double)x + dx;
            if (sy < -0.5 || sy > rows_f - 0.5 || sx < -0.5 || sx > cols_f - 0.5) continue;
            row[x] = orc_bicubic_sample(img, rows, cols, sy, sx);
        }
    }
}

int orc_resample_image(const float *img, size_t src_rows, size_t src_cols, size_t target_rows,
                       size_t target_cols, float *out) {
    if (target_rows == 0 || target_cols == 0) return 1;  // bail!
    if (target_rows == src_rows && target_cols == src_cols) {
        std::memcpy(out, img, src_rows * src_cols * sizeof(float));
        return 0;
    }
    const double scale_y = (double)src_rows / (double)target_rows;
    const double scale_x = (double)src_cols / (double)target_cols;
    const double half_shift_y = (scale_y - 1.0) * 0.5;
    const double half_shift_x = (scale_x - 1.0) * 0.5;
#pragma omp parallel for schedule(dynamic, 8)
    for (long long ty = 0; ty < (long long)target_rows; ++ty) {
        const double sy = (double)ty * scale_y + half_shift_y;
        float *row = out + (size_t)ty * target_cols;
        for (size_t tx = 0; tx < target_cols; ++tx) {
            const double sx = (double)tx * scale_x + half_shift_x;
            row[tx] = orc_bicubic_sample(img, src_rows, src_cols, sy, sx);
        }
    }
    return 0;
}

}  // extern "C"
