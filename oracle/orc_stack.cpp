// orc_stack.cpp -- CPU oracle (TEST INFRASTRUCTURE, see oracle.h) for the
// sigma-clip frame stack and the order-statistic helpers it shares with the
// statistics and background code.
//
// Restates (paths relative to src-tauri/src/):
//   core/stacking/combine.rs:14-92   sigma_clip_combine
//   core/stacking/combine.rs:94-193  stack_images (align=false core loop)
//   math/median.rs:4-73              f32_cmp, exact_median_mut, median_f32_mut, exact_mad_mut
//   math/sigma_clip.rs:4-34          sigma_clipped_stats
//
// Rust std semantics mirrored: select_nth_unstable_by == std::nth_element for
// the element at k (the permutation of the rest differs, which only changes
// the order of later f64 sums); f64::max ignores NaN (std::fmax); `as f32`
// rounds to nearest even.  Build with -ffp-contract=off: Rust never fuses.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include <omp.h>

#include "oracle.h"

namespace {

constexpr double MAD_TO_SIGMA = 1.4826;  // types/constants.rs:7

// math/median.rs:4-13 -- total order with NaN last.
inline bool f32_less(float a, float b) {
    const bool an = std::isnan(a), bn = std::isnan(b);
    if (an || bn) return !an && bn;
    return a < b;
}

}  // namespace

extern "C" {

int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
}

// math/median.rs:27-44
double orc_exact_median_mut(float *data, size_t n) {
    if (n == 0) return 0.0;
    const size_t mid = n / 2;
    std::nth_element(data, data + mid, data + n, f32_less);
    if (n % 2 == 0) {
        const double right = (double)data[mid];
        float left = -FLT_MAX;  // f32::MIN
        for (size_t i = 0; i < mid; ++i)
            if (data[i] > left) left = data[i];
        return ((double)left + right) / 2.0;
    }
    return (double)data[mid];
}

// math/median.rs:46-63
float orc_median_f32_mut(float *data, size_t n) {
    if (n == 0) return 0.0f;
    const size_t mid = n / 2;
    std::nth_element(data, data + mid, data + n, f32_less);
    if (n % 2 == 0) {
        const float right = data[mid];
        float left = -FLT_MAX;
        for (size_t i = 0; i < mid; ++i)
            if (data[i] > left) left = data[i];
        return (left + right) / 2.0f;
    }
    return data[mid];
}

// math/median.rs:65-73
float orc_exact_mad_mut(float *data, size_t n, float median) {
    if (n == 0) return 0.0f;
    for (size_t i = 0; i < n; ++i) data[i] = std::fabs(data[i] - median);
    return orc_median_f32_mut(data, n);
}

// math/sigma_clip.rs:4-34
void orc_sigma_clipped_stats(float *values, size_t *n_io, float kappa, size_t iterations,
                             double *median_out, double *sigma_out) {
    size_t n = *n_io;
    std::vector<float> devs;
    devs.reserve(n);
    for (size_t it = 0; it < iterations; ++it) {
        if (n < 3) break;
        const double median = orc_exact_median_mut(values, n);
        devs.resize(n);
        for (size_t i = 0; i < n; ++i) devs[i] = (float)std::fabs((double)values[i] - median);
        const double mad = (double)orc_median_f32_mut(devs.data(), n);
        const double sig = std::fmax(mad * MAD_TO_SIGMA, 1e-30);
        const float lo = (float)(median - (double)kappa * sig);
        const float hi = (float)(median + (double)kappa * sig);
        size_t w = 0;
        for (size_t i = 0; i < n; ++i) {  // Vec::retain keeps order
            const float v = values[i];
            if (v >= lo && v <= hi) values[w++] = v;
        }
        n = w;
    }
    *n_io = n;
    if (n == 0) {
        *median_out = 0.0;
        *sigma_out = 1.0;
        return;
    }
    const double median = orc_exact_median_mut(values, n);
    devs.resize(n);
    for (size_t i = 0; i < n; ++i) devs[i] = (float)std::fabs((double)values[i] - median);
    *median_out = median;
    *sigma_out = std::fmax((double)orc_median_f32_mut(devs.data(), n) * MAD_TO_SIGMA, 1e-30);
}

// core/stacking/combine.rs:14-92
float orc_sigma_clip_combine(float *values, size_t n_orig, float sigma_low, float sigma_high,
                             size_t max_iter, uint32_t *rejected_out) {
    *rejected_out = 0;
    if (n_orig == 0) return 0.0f;
    if (n_orig == 1) return values[0];

    size_t len = n_orig;
    uint32_t rejected = 0;
    float last_center = NAN;
    std::vector<float> devs;

    for (size_t iteration = 0; iteration < max_iter; ++iteration) {
        if (len < 2) break;
        float center, sigma;
        if (iteration == 0) {
            // upper median, NOT averaged for even len (combine.rs:38-40)
            const size_t mid = len / 2;
            std::nth_element(values, values + mid, values + len, f32_less);
            const float med = values[mid];
            devs.resize(len);
            for (size_t i = 0; i < len; ++i) devs[i] = std::fabs(values[i] - med);
            const size_t dmid = len / 2;
            std::nth_element(devs.begin(), devs.begin() + dmid, devs.end(), f32_less);
            const float mad = devs[dmid];
            center = med;
            sigma = (float)std::fmax((double)mad * MAD_TO_SIGMA, 1e-10);
        } else {
            const double n = (double)len;
            double sum = 0.0;
            for (size_t i = 0; i < len; ++i) sum += (double)values[i];
            const double mean = sum / n;
            double ss = 0.0;
            for (size_t i = 0; i < len; ++i) {
                const double d = (double)values[i] - mean;
                ss += d * d;
            }
            const double variance = ss / std::fmax(n - 1.0, 1.0);
            center = (float)mean;
            sigma = (float)std::fmax(std::sqrt(variance), 1e-10);
        }
        last_center = center;

        const float lo = -sigma_low * sigma;
        const float hi = sigma_high * sigma;
        size_t write = 0;
        for (size_t read = 0; read < len; ++read) {
            const float dev = values[read] - center;
            if (dev >= lo && dev <= hi) values[write++] = values[read];
        }
        const size_t removed = len - write;
        rejected += (uint32_t)removed;
        len = write;
        if (removed == 0) break;
    }

    *rejected_out = rejected;
    if (len == 0) return std::isfinite(last_center) ? last_center : 0.0f;

    double sum = 0.0;
    for (size_t i = 0; i < len; ++i) sum += (double)values[i];
    return (float)(sum / (double)len);
}

// core/stacking/combine.rs:94-193 with config.align == false; frames are
// already cropped to rows x cols by the caller (combine.rs:104-113 is a copy).
int orc_stack_images(const float *const *frames, size_t n_frames, size_t rows, size_t cols,
                     float sigma_low, float sigma_high, size_t max_iter, float *out,
                     uint64_t *rejected_pixels) {
    if (n_frames == 0) return 1;  // bail!("No images to stack")
    uint64_t total_rejected = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : total_rejected)
    for (long long y = 0; y < (long long)rows; ++y) {
        std::vector<float> vals(n_frames);
        const size_t base = (size_t)y * cols;
        uint64_t local = 0;
        for (size_t x = 0; x < cols; ++x) {
            size_t cnt = 0;
            for (size_t f = 0; f < n_frames; ++f) {
                const float v = frames[f][base + x];
                if (std::isfinite(v)) vals[cnt++] = v;
            }
            uint32_t rej = 0;
            out[base + x] =
                orc_sigma_clip_combine(vals.data(), cnt, sigma_low, sigma_high, max_iter, &rej);
            local += rej;
        }
        total_rejected += local;
    }
    *rejected_pixels = total_rejected;
    return 0;
}

}  // extern "C"
