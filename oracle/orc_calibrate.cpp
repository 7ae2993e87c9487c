// orc_calibrate.cpp -- CPU oracle (TEST INFRASTRUCTURE, see oracle.h) for the calibration
// masters and frame calibration.
//
// Restates src-tauri/src/core/stacking/calibration.rs:
//   :15-45   subtract_bias, subtract_dark, divide_flat (calibrate without the final clamp)
//   :47-83   calibrate_image
//   :85-123  median_combine_row_major (upper median vals[len/2] of the finite samples; empty -> 0)
//   :160-186 create_master_dark's preprocessing (frame - bias)
//   :199-229 create_master_flat's preprocessing (frame - bias, then - dark * 1.0)
//   :231-252 master-flat normalisation (sequential f64 sum of the finite positive pixels)
// and src-tauri/src/core/imaging/calibration_pipeline.rs:
//   :317-378 sigma_clipped_mean_stack (the batch pipeline's clip: MAD every iteration, strict
//            z-score test, per-frame rejection counts, f32 mean in frame order)
// FITS loading (load_fits_image) is outside the path: frames arrive as arrays.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

#include "oracle.h"

namespace {
// math/median.rs:4-13 -- total order with NaN last (only finite values reach it here)
inline bool f32_less(float a, float b) {
    const bool an = std::isnan(a), bn = std::isnan(b);
    if (an || bn) return !an && bn;
    return a < b;
}
}  // namespace

extern "C" {

int orc_median_combine(const float *const *frames, size_t n_frames, size_t rows, size_t cols,
                       const float *bias, const float *dark, float dark_ratio, float *out) {
    if (n_frames == 0) return 1;  // "No ... frames provided", calibration.rs:126-128
#pragma omp parallel for schedule(dynamic)
    for (long long y = 0; y < (long long)rows; ++y) {
        std::vector<float> vals;
        vals.reserve(n_frames);
        const size_t base = (size_t)y * cols;
        for (size_t x = 0; x < cols; ++x) {
            vals.clear();
            const size_t idx = base + x;
            for (size_t f = 0; f < n_frames; ++f) {
                float v = frames[f][idx];
                if (bias) v = v - bias[idx];               // subtract_bias, :15-17
                if (dark) v = v - dark[idx] * dark_ratio;  // subtract_dark, :19-25
                if (std::isfinite(v)) vals.push_back(v);
            }
            if (vals.empty()) {
                out[idx] = 0.0f;
            } else {
                const size_t mid = vals.size() / 2;
                std::nth_element(vals.begin(), vals.begin() + mid, vals.end(), f32_less);
                out[idx] = vals[mid];
            }
        }
    }
    return 0;
}

// calibration_pipeline.rs:317-378.  rejection_counts: n_frames entries, overwritten.
int orc_sigma_clipped_mean_stack(const float *const *frames, size_t n_frames, size_t rows, size_t cols,
                                 float sigma_low, float sigma_high, size_t max_iterations, float *out,
                                 uint64_t *rejection_counts) {
    if (n_frames == 0) return 1;
    const double MAD_TO_SIGMA = 1.4826;  // types/constants.rs:7
    std::vector<uint64_t> total(n_frames, 0);
#pragma omp parallel
    {
        std::vector<uint64_t> local(n_frames, 0);
        std::vector<std::pair<float, size_t>> vals;
        std::vector<float> scratch;
        vals.reserve(n_frames);
        scratch.reserve(n_frames);
#pragma omp for schedule(dynamic)
        for (long long y = 0; y < (long long)rows; ++y) {
            const size_t base = (size_t)y * cols;
            for (size_t x = 0; x < cols; ++x) {
                vals.clear();
                const size_t idx = base + x;
                for (size_t i = 0; i < n_frames; ++i) vals.push_back({frames[i][idx], i});
                for (size_t it = 0; it < max_iterations; ++it) {
                    if (vals.size() < 3) break;
                    scratch.clear();
                    for (const auto &v : vals) scratch.push_back(v.first);
                    const size_t mid = scratch.size() / 2;
                    std::nth_element(scratch.begin(), scratch.begin() + mid, scratch.end(), f32_less);
                    const float median = scratch[mid];
                    for (float &v : scratch) v = std::fabs(v - median);
                    std::nth_element(scratch.begin(), scratch.begin() + mid, scratch.end(), f32_less);
                    const float sigma = (float)((double)scratch[mid] * MAD_TO_SIGMA);
                    if (sigma < 1e-10f) break;
                    const size_t before = vals.size();
                    size_t w = 0;
                    for (size_t r = 0; r < vals.size(); ++r) {  // Vec::retain keeps order
                        const float z = (vals[r].first - median) / sigma;
                        const bool keep = z > -sigma_low && z < sigma_high;
                        if (!keep) local[vals[r].second] += 1;
                        else vals[w++] = vals[r];
                    }
                    vals.resize(w);
                    if (vals.size() == before) break;
                }
                if (vals.empty()) {
                    out[idx] = 0.0f;
                } else {
                    float sum = 0.0f;  // iter().sum::<f32>(): sequential, frame order
                    for (const auto &v : vals) sum += v.first;
                    out[idx] = sum / (float)vals.size();
                }
            }
        }
#pragma omp critical
        for (size_t i = 0; i < n_frames; ++i) total[i] += local[i];
    }
    for (size_t i = 0; i < n_frames; ++i) rejection_counts[i] = total[i];
    return 0;
}

void orc_normalize_flat(float *data, size_t len) {
    double sum = 0.0;
    size_t count = 0;
    for (size_t i = 0; i < len; ++i) {
        const float v = data[i];
        if (std::isfinite(v) && v > 0.0f) {
            sum += (double)v;
            ++count;
        }
    }
    if (count == 0) return;
    const double mean = sum / (double)count;
    const float inv_mean = std::fabs(mean) > 1e-10 ? 1.0f / (float)mean : 1.0f;
    for (size_t i = 0; i < len; ++i) {
        float &v = data[i];
        if (std::isfinite(v) && v > 0.0f) v *= inv_mean;
        else v = 1.0f;
    }
}

void orc_calibrate_image(const float *raw, size_t len, const float *bias, const float *dark,
                         const float *flat, float dark_ratio, int clamp_negative, float *out) {
#pragma omp parallel for
    for (long long i = 0; i < (long long)len; ++i) {
        float v = raw[i];
        if (bias) v -= bias[i];
        if (dark) v -= dark[i] * dark_ratio;
        if (flat) {
            const float fv = flat[i];
            if (std::isfinite(fv) && std::fabs(fv) > 1e-4f) v /= fv;
        }
        out[i] = (clamp_negative && v < 0.0f) ? 0.0f : v;
    }
}

}  // extern "C"
