// orc_preview.cpp -- CPU oracle (TEST INFRASTRUCTURE, see oracle.h) for the preview/export tail
// and the masked stretch.
//
// Restates (paths relative to src-tauri/src/):
//   infra/render/rgb.rs:7-47,49-98     render_rgb (u8), render_rgb_16bit (u16 big-endian)
//   cmd/helpers.rs:204-262             render_rgb_preview (nearest decimation to max_dim)
//   cmd/helpers.rs:264-322             render_rgb_preview_with_stf (per-channel STF -> u8)
//   core/imaging/stf.rs:122-145        make_stf_u8_fn
//   core/compose/lrgb.rs:4-62          apply_lrgb, synthesize_luminance
//   core/imaging/star_mask.rs:46-138   generate_star_mask_from_detection
//   core/imaging/masked_stretch.rs:59-258  masked_stretch_with_mask, compute_luminance,
//                                      normalize_to_01, compute_masked_median, mtf_balance, apply_mtf
//
// PARITY PINNING: the reference holds unit tests for lrgb.rs only (ported in
// tests/test_oracle_kat.py::test_lrgb_reference_unit_tests).  star_mask.rs, masked_stretch.rs and
// the render helpers have no tests and no golden vectors: for those this restatement is checked
// against numpy restatements of their closed-form parts and is otherwise "parity unpinned".
//
// Build with -ffp-contract=off (Rust never fuses mul+add).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

inline float clampf(float x, float lo, float hi) {  // f32::clamp (NaN stays NaN)
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}
inline double clampd(double x, double lo, double hi) {
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}
// Rust `f32 as u8` / `as u16`: truncating, saturating, NaN -> 0
inline unsigned sat_trunc(float x, unsigned hi) {
    if (!(x == x)) return 0;
    if (x <= 0.0f) return 0;
    if (x >= (float)hi) return hi;
    return (unsigned)x;
}
// Rust `f64 as usize`
inline size_t sat_usize(double x) {
    if (!(x == x) || x <= 0.0) return 0;
    if (x >= 18446744073709551615.0) return (size_t)-1;
    return (size_t)x;
}
inline double mtf(double x, double m) {  // stf.rs:50-58
    if (x <= 0.0) return 0.0;
    if (x >= 1.0) return 1.0;
    return (m - 1.0) * x / ((2.0 * m - 1.0) * x - m);
}
inline bool valid_px(float v) { return std::isfinite(v) && v > 1e-7f; }

struct StfU8 {  // stf.rs:122-145
    double inv_range, dmin, shadow, inv_clip, midtone;
    StfU8(const orc_stf_params &p, const orc_image_stats &st) {
        inv_range = 1.0 / std::fmax(st.max - st.min, 1e-30);
        dmin = st.min;
        shadow = p.shadow;
        inv_clip = 1.0 / std::fmax(p.highlight - p.shadow, 1e-15);
        midtone = p.midtone;
    }
    uint8_t operator()(float v) const {
        if (!valid_px(v)) return 0;
        const double norm = ((double)v - dmin) * inv_range;
        const double clipped = clampd((norm - shadow) * inv_clip, 0.0, 1.0);
        const double y = std::round(mtf(clipped, midtone) * 255.0);  // half away from zero
        if (!(y == y) || y <= 0.0) return 0;
        return y >= 255.0 ? 255 : (uint8_t)y;
    }
};

}  // namespace

extern "C" {

// cmd/helpers.rs:222-227 / :282-291: preview size; max_dim == 0 means "no decimation"
void orc_preview_size(size_t rows, size_t cols, size_t max_dim, size_t *pw, size_t *ph) {
    if (max_dim == 0 || (rows <= max_dim && cols <= max_dim)) {
        *pw = cols;
        *ph = rows;
        return;
    }
    const double scale = (double)max_dim / (double)std::max(rows, cols);
    *pw = sat_usize(std::fmax(std::round((double)cols * scale), 1.0));
    *ph = sat_usize(std::fmax(std::round((double)rows * scale), 1.0));
}

// render_rgb (rgb.rs:7-47) when the planes fit max_dim, else render_rgb_preview (helpers.rs:204-262)
void orc_render_rgb_u8(const float *r, const float *g, const float *b, size_t rows, size_t cols,
                       size_t max_dim, uint8_t *out) {
    size_t pw, ph;
    orc_preview_size(rows, cols, max_dim, &pw, &ph);
    const bool full = pw == cols && ph == rows && (max_dim == 0 || (rows <= max_dim && cols <= max_dim));
    const double y_ratio = (double)rows / (double)ph, x_ratio = (double)cols / (double)pw;
#pragma omp parallel for schedule(static)
    for (long long dy = 0; dy < (long long)ph; ++dy) {
        const size_t sy = full ? (size_t)dy : sat_usize(std::fmin((double)dy * y_ratio, (double)(rows - 1)));
        for (size_t dx = 0; dx < pw; ++dx) {
            const size_t sx = full ? dx : sat_usize(std::fmin((double)dx * x_ratio, (double)(cols - 1)));
            const size_t si = sy * cols + sx;
            uint8_t *o = out + ((size_t)dy * pw + dx) * 3;
            o[0] = (uint8_t)sat_trunc(clampf(r[si], 0.0f, 1.0f) * 255.0f, 255);
            o[1] = (uint8_t)sat_trunc(clampf(g[si], 0.0f, 1.0f) * 255.0f, 255);
            o[2] = (uint8_t)sat_trunc(clampf(b[si], 0.0f, 1.0f) * 255.0f, 255);
        }
    }
}

// render_rgb_16bit, rgb.rs:49-98: (v.clamp(0,1) * 65535) as u16, big-endian bytes, interleaved
void orc_render_rgb_u16be(const float *r, const float *g, const float *b, size_t len, uint8_t *out) {
    const float *ch[3] = {r, g, b};
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)len; ++i)
        for (int c = 0; c < 3; ++c) {
            const unsigned v = sat_trunc(clampf(ch[c][i], 0.0f, 1.0f) * 65535.0f, 65535);
            out[(size_t)i * 6 + c * 2] = (uint8_t)(v >> 8);
            out[(size_t)i * 6 + c * 2 + 1] = (uint8_t)(v & 255);
        }
}

// render_rgb_preview_with_stf, helpers.rs:264-322 (always the ratio path; ratios are 1.0 when it fits)
void orc_render_rgb_stf_u8(const float *r, const float *g, const float *b, size_t rows, size_t cols,
                           const orc_stf_params *stf3, const orc_image_stats *st3, size_t max_dim,
                           uint8_t *out) {
    size_t pw, ph;
    orc_preview_size(rows, cols, max_dim, &pw, &ph);
    const bool fits = max_dim == 0 || (rows <= max_dim && cols <= max_dim);
    const double y_ratio = fits ? 1.0 : (double)rows / (double)ph;
    const double x_ratio = fits ? 1.0 : (double)cols / (double)pw;
    const StfU8 fr(stf3[0], st3[0]), fg(stf3[1], st3[1]), fb(stf3[2], st3[2]);
#pragma omp parallel for schedule(static)
    for (long long dy = 0; dy < (long long)ph; ++dy) {
        const size_t sy = sat_usize(std::fmin((double)dy * y_ratio, (double)(rows - 1)));
        for (size_t dx = 0; dx < pw; ++dx) {
            const size_t sx = sat_usize(std::fmin((double)dx * x_ratio, (double)(cols - 1)));
            const size_t si = sy * cols + sx;
            uint8_t *o = out + ((size_t)dy * pw + dx) * 3;
            o[0] = fr(r[si]);
            o[1] = fg(g[si]);
            o[2] = fb(b[si]);
        }
    }
}

// apply_lrgb, lrgb.rs:4-46 (dimension check is the caller's)
void orc_apply_lrgb(const float *l, float *r, float *g, float *b, size_t len, float lightness_weight,
                    float chrominance_weight) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)len; ++i) {
        const float lum_new = l[i];
        const float lum_old = r[i] * 0.2126f + g[i] * 0.7152f + b[i] * 0.0722f;
        if (lum_old < 1e-10f) {
            const float blended = lum_new * lightness_weight;
            r[i] = g[i] = b[i] = blended;
            continue;
        }
        const float ratio = (lum_new * lightness_weight + lum_old * (1.0f - lightness_weight)) / lum_old;
        const float cb = chrominance_weight;
        r[i] = clampf(r[i] * ratio * cb + lum_new * (1.0f - cb), 0.0f, 1.0f);
        g[i] = clampf(g[i] * ratio * cb + lum_new * (1.0f - cb), 0.0f, 1.0f);
        b[i] = clampf(b[i] * ratio * cb + lum_new * (1.0f - cb), 0.0f, 1.0f);
    }
}

// synthesize_luminance, lrgb.rs:48-62 (finite_only = 0); compute_luminance,
// masked_stretch.rs:134-159 (finite_only = 1: non-finite samples count as 0, weights first)
void orc_luminance(const float *r, const float *g, const float *b, size_t len, int finite_only,
                   float *out) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)len; ++i) {
        if (finite_only) {
            const float rn = std::isfinite(r[i]) ? r[i] : 0.0f;
            const float gn = std::isfinite(g[i]) ? g[i] : 0.0f;
            const float bn = std::isfinite(b[i]) ? b[i] : 0.0f;
            out[i] = 0.2126f * rn + 0.7152f * gn + 0.0722f * bn;
        } else {
            out[i] = r[i] * 0.2126f + g[i] * 0.7152f + b[i] * 0.0722f;
        }
    }
}

// generate_star_mask_from_detection, star_mask.rs:46-138.  Returns stars_masked.
size_t orc_star_mask(const float *img, size_t h, size_t w, const orc_star *stars, size_t n_stars,
                     double growth_factor, double softness, double min_fwhm, double max_fwhm,
                     int luminance_protect, double luminance_ceiling, float *mask,
                     double *coverage_fraction) {
    std::fill(mask, mask + h * w, 0.0f);
    size_t count = 0;
    for (size_t s = 0; s < n_stars; ++s) {
        const orc_star &st = stars[s];
        if (!(st.fwhm >= min_fwhm && st.fwhm <= max_fwhm)) continue;
        ++count;
        const double radius = st.fwhm * growth_factor;
        const double soft_radius = radius + softness;
        const size_t y_min = sat_usize(std::fmax(std::floor(st.y - soft_radius), 0.0));
        const size_t y_max = std::min(sat_usize(std::ceil(st.y + soft_radius)), h ? h - 1 : 0);
        const size_t x_min = sat_usize(std::fmax(std::floor(st.x - soft_radius), 0.0));
        const size_t x_max = std::min(sat_usize(std::ceil(st.x + soft_radius)), w ? w - 1 : 0);
        const double r2_inner = radius * radius;
        const double r2_outer = soft_radius * soft_radius;
        const double fade_range = std::fmax(r2_outer - r2_inner, 1e-10);
        for (size_t py = y_min; py <= y_max && py < h; ++py)
            for (size_t px = x_min; px <= x_max && px < w; ++px) {
                const double dx = (double)px - st.x, dy = (double)py - st.y;
                const double d2 = dx * dx + dy * dy;
                float val;
                if (d2 <= r2_inner) {
                    val = 1.0f;
                } else if (d2 <= r2_outer) {
                    const float t = (float)((d2 - r2_inner) / fade_range);
                    const float smooth = t * t * (3.0f - 2.0f * t);
                    val = 1.0f - smooth;
                } else {
                    continue;
                }
                float &m = mask[py * w + px];
                if (val > m) m = val;
            }
    }
    if (luminance_protect) {
        const float ceiling = (float)luminance_ceiling;
        const float inv_range = ceiling < 1.0f ? 1.0f / (1.0f - ceiling) : 1.0f;
        for (size_t i = 0; i < h * w; ++i) {
            const float pixel = img[i];
            if (pixel > ceiling && mask[i] < 1.0f) {
                const float excess = clampf((pixel - ceiling) * inv_range, 0.0f, 1.0f);
                const float smooth = excess * excess * (3.0f - 2.0f * excess);
                if (smooth > mask[i]) mask[i] = smooth;
            }
        }
    }
    size_t covered = 0;
    for (size_t i = 0; i < h * w; ++i) covered += mask[i] > 0.01f;
    *coverage_fraction = (double)covered / (double)(h * w);
    return count;
}

// compute_masked_median, masked_stretch.rs:211-228
double orc_masked_median(const float *img, const float *mask, size_t len) {
    std::vector<float> v;
    v.reserve(len);
    for (size_t i = 0; i < len; ++i)
        if (mask[i] < 0.5f && std::isfinite(img[i]) && img[i] > 0.0f) v.push_back(img[i]);
    if (v.empty()) return 0.0;
    const size_t mid = v.size() / 2;
    std::nth_element(v.begin(), v.begin() + mid, v.end());
    return (double)v[mid];
}

// masked_stretch_with_mask, masked_stretch.rs:59-117 (+ normalize_to_01 :192-209, mtf_balance
// :230-236, apply_mtf :238-254, clamp_inplace :256-258).  Returns iterations_run.
size_t orc_masked_stretch_with_mask(const float *img, size_t len, const float *mask,
                                    size_t iterations, double target_bg, double protection_amount,
                                    double convergence_threshold, float *out,
                                    double *final_background, int *converged) {
    orc_image_stats st;
    orc_compute_image_stats(img, len, &st);
    const float range = (float)(st.max - st.min);
    if (range < 1e-10f) {
        std::fill(out, out + len, 0.0f);
    } else {
        const float dmin = (float)st.min;
        const float inv = 1.0f / range;
        for (size_t i = 0; i < len; ++i) {
            const float v = img[i];
            out[i] = (!std::isfinite(v) || v <= 0.0f) ? 0.0f : clampf((v - dmin) * inv, 0.0f, 1.0f);
        }
    }
    const float protection = (float)protection_amount;
    double prev_bg = orc_masked_median(out, mask, len);
    size_t iterations_run = 0;
    *converged = 0;
    for (size_t it = 0; it < iterations; ++it) {
        iterations_run = it + 1;
        const double bg = orc_masked_median(out, mask, len);
        const bool at_target = std::fabs(bg - target_bg) < convergence_threshold;
        const bool stagnated = it > 0 && std::fabs(bg - prev_bg) < convergence_threshold * 0.1;
        if (at_target) {
            *converged = 1;
            break;
        }
        if (stagnated) break;
        const double denom = 2.0 * target_bg * bg - target_bg - bg;
        const double midtone =
            std::fabs(denom) < 1e-15 ? 0.5 : clampd(bg * (target_bg - 1.0) / denom, 0.0001, 0.9999);
        const float m = (float)midtone;
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)len; ++i) {
            const float x = out[i];
            float stretched;
            if (x <= 0.0f) stretched = 0.0f;
            else if (x >= 1.0f) stretched = 1.0f;
            else {
                const float d = (2.0f * m - 1.0f) * x - m;
                stretched = std::fabs(d) < 1e-10f ? x : clampf((m - 1.0f) * x / d, 0.0f, 1.0f);
            }
            const float blend = mask[i] * protection;
            out[i] = x * blend + stretched * (1.0f - blend);
        }
        prev_bg = bg;
    }
    *final_background = orc_masked_median(out, mask, len);
    for (size_t i = 0; i < len; ++i) out[i] = clampf(out[i], 0.0f, 1.0f);
    return iterations_run;
}

}  // extern "C"
