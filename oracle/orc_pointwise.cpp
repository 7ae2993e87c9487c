// orc_pointwise.cpp -- CPU oracle (TEST INFRASTRUCTURE, see oracle.h) for the
// one-sweep per-pixel operators.
//
// Restates (paths relative to src-tauri/src/):
//   core/imaging/stf.rs:50-120        mtf, StfTransform, apply_stf, apply_stf_f32
//   core/compose/rgb.rs:191-207       apply_stf_inplace (divide-by-clip_range twin)
//   core/imaging/stretch.rs:10-46     arcsinh_stretch_with_stats
//   core/compose/channel_blend.rs:13-70  blend_channels
//   core/imaging/scnr.rs:7-53         apply_scnr_inplace
//   core/compose/rgb.rs:127-130, cmd/compose/color.rs:21-49   v * factor
//   core/imaging/curves.rs:17-63      levels
//   core/imaging/curves.rs:65-208     SplineLut, Fritsch-Carlson, apply_curve
//
// Build with -ffp-contract=off (Rust never fuses mul+add).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

constexpr float PADDING_THRESHOLD = 1e-7f;
inline bool valid_px(float v) { return std::isfinite(v) && v > PADDING_THRESHOLD; }

inline double clampd(double x, double lo, double hi) {
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}
inline float clampf(float x, float lo, float hi) {
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}

// stf.rs:50-58
inline double mtf(double x, double m) {
    if (x <= 0.0) return 0.0;
    if (x >= 1.0) return 1.0;
    return (m - 1.0) * x / ((2.0 * m - 1.0) * x - m);
}

struct StfTransform {  // stf.rs:60-87
    double inv_range, dmin, shadow, inv_clip, midtone;
    StfTransform(const orc_stf_params *p, const orc_image_stats *st) {
        const double range = std::fmax(st->max - st->min, 1e-30);
        const double clip_range = std::fmax(p->highlight - p->shadow, 1e-15);
        inv_range = 1.0 / range;
        dmin = st->min;
        shadow = p->shadow;
        inv_clip = 1.0 / clip_range;
        midtone = p->midtone;
    }
    inline double apply(double v) const {
        const double norm = (v - dmin) * inv_range;  // NOT clamped (G6)
        const double clipped = clampd((norm - shadow) * inv_clip, 0.0, 1.0);
        return mtf(clipped, midtone);
    }
};

// Rust f32::asinh (library/std/src/f32.rs): ax + ax/(hypot(1, 1/ax) + 1/ax), ln_1p, copysign.
inline float rust_asinhf(float x) {
    const float ax = std::fabs(x);
    const float ix = 1.0f / ax;
    return std::copysign(std::log1p(ax + (ax / (std::hypot(1.0f, ix) + ix))), x);
}

// Rust `f32 as usize` for the LUT index.
inline size_t f2us(float x) {
    if (!(x > 0.0f)) return 0;
    return (size_t)x;
}

// curves.rs:113-151
std::vector<double> fritsch_carlson_tangents(const std::vector<std::pair<double, double>> &pts) {
    const size_t n = pts.size();
    if (n < 2) return std::vector<double>(n, 0.0);
    if (n == 2) {
        const double slope =
            (pts[1].second - pts[0].second) / std::fmax(pts[1].first - pts[0].first, 1e-15);
        return {slope, slope};
    }
    std::vector<double> deltas, slopes;
    for (size_t i = 0; i + 1 < n; ++i) {
        const double dx = std::fmax(pts[i + 1].first - pts[i].first, 1e-15);
        deltas.push_back(dx);
        slopes.push_back((pts[i + 1].second - pts[i].second) / dx);
    }
    // f64::signum: +1 for +0.0/positive, -1 for -0.0/negative, NaN for NaN
    auto signum = [](double x) {
        if (std::isnan(x)) return x;
        return std::signbit(x) ? -1.0 : 1.0;
    };
    std::vector<double> m(n, 0.0);
    m[0] = slopes[0];
    m[n - 1] = slopes[n - 2];
    for (size_t i = 1; i + 1 < n; ++i) {
        if (signum(slopes[i - 1]) != signum(slopes[i]))
            m[i] = 0.0;
        else
            m[i] = (slopes[i - 1] + slopes[i]) * 0.5;
    }
    for (size_t i = 0; i + 1 < n; ++i) {
        if (std::fabs(slopes[i]) < 1e-15) {
            m[i] = 0.0;
            m[i + 1] = 0.0;
            continue;
        }
        const double alpha = m[i] / slopes[i];
        const double beta = m[i + 1] / slopes[i];
        const double tau = alpha * alpha + beta * beta;
        if (tau > 9.0) {
            const double s = 3.0 / std::sqrt(tau);
            m[i] = s * alpha * slopes[i];
            m[i + 1] = s * beta * slopes[i];
        }
    }
    return m;
}

// curves.rs:153-184
double hermite_eval(const std::vector<std::pair<double, double>> &pts,
                    const std::vector<double> &tan, size_t n, double x) {
    if (x <= pts[0].first) return pts[0].second;
    if (x >= pts[n - 1].first) return pts[n - 1].second;
    size_t seg = 0;
    for (size_t i = 1; i < n; ++i) {
        if (x < pts[i].first) {
            seg = i - 1;
            break;
        }
    }
    const double dx = std::fmax(pts[seg + 1].first - pts[seg].first, 1e-15);
    const double t = (x - pts[seg].first) / dx;
    const double t2 = t * t;
    const double t3 = t2 * t;
    const double h00 = 2.0 * t3 - 3.0 * t2 + 1.0;
    const double h10 = t3 - 2.0 * t2 + t;
    const double h01 = -2.0 * t3 + 3.0 * t2;
    const double h11 = t3 - t2;
    return h00 * pts[seg].second + h10 * dx * tan[seg] + h01 * pts[seg + 1].second +
           h11 * dx * tan[seg + 1];
}

}  // namespace

extern "C" {

// stf.rs:89-102
void orc_apply_stf_u8(const float *data, size_t len, const orc_stf_params *p,
                      const orc_image_stats *st, uint8_t *out) {
    const StfTransform tx(p, st);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)len; ++i) {
        const float v = data[i];
        if (!valid_px(v)) {
            out[i] = 0;
            continue;
        }
        const double y = clampd(std::round(tx.apply((double)v) * 255.0), 0.0, 255.0);
        out[i] = std::isnan(y) ? 0 : (uint8_t)y;
    }
}

// stf.rs:104-120
void orc_apply_stf_f32(const float *data, size_t len, const orc_stf_params *p,
                       const orc_image_stats *st, float *out) {
    const StfTransform tx(p, st);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)len; ++i) {
        const float v = data[i];
        out[i] = valid_px(v) ? (float)tx.apply((double)v) : 0.0f;
    }
}

// compose/rgb.rs:191-207
void orc_apply_stf_rgbvariant_f32(const float *data, size_t len, const orc_stf_params *p,
                                  const orc_image_stats *st, float *out) {
    const double range = std::fmax(st->max - st->min, 1e-30);
    const double inv_range = 1.0 / range;
    const double dmin = st->min, shadow = p->shadow, highlight = p->highlight;
    const double clip_range = std::fmax(highlight - shadow, 1e-15);
    const double m = p->midtone;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)len; ++i) {
        const float v = data[i];
        if (!std::isfinite(v) || v <= 1e-7f) {
            out[i] = 0.0f;
            continue;
        }
        const double norm = ((double)v - dmin) * inv_range;
        const double clipped = clampd((norm - shadow) / clip_range, 0.0, 1.0);
        if (clipped <= 0.0)
            out[i] = 0.0f;
        else if (clipped >= 1.0)
            out[i] = 1.0f;
        else
            out[i] = (float)((m - 1.0) * clipped / ((2.0 * m - 1.0) * clipped - m));
    }
}

// stretch.rs:10-46
void orc_arcsinh_stretch(const float *data, size_t len, float dmin, float dmax, float factor,
                         float gamma, float *out) {
    if (std::fabs(factor) < 1e-10f) {
        std::memcpy(out, data, len * sizeof(float));
        return;
    }
    const float range = dmax - dmin;
    if (range < 1e-10f) {
        std::memset(out, 0, len * sizeof(float));
        return;
    }
    const float inv_range = 1.0f / range;
    const float inv_denom = 1.0f / rust_asinhf(factor);
    const bool apply_gamma = std::fabs(gamma - 1.0f) > 1e-6f;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)len; ++i) {
        const float val = data[i];
        if (!std::isfinite(val)) {
            out[i] = 0.0f;
        } else {
            const float norm = clampf((val - dmin) * inv_range, 0.0f, 1.0f);
            const float stretched = rust_asinhf(norm * factor) * inv_denom;
            out[i] = apply_gamma ? std::pow(stretched, gamma) : stretched;
        }
    }
}

// channel_blend.rs:13-70
void orc_blend_channels(const float *const *channels, const size_t *channel_lens, size_t n_ch,
                        const uint64_t *w_idx, const double *w_rgb, size_t n_w, size_t rows,
                        size_t cols, float *r, float *g, float *b) {
    struct W {
        size_t ch;
        float rw, gw, bw;
    };
    std::vector<W> vw;
    for (size_t k = 0; k < n_w; ++k)
        if (w_idx[k] < n_ch)
            vw.push_back({(size_t)w_idx[k], (float)w_rgb[3 * k], (float)w_rgb[3 * k + 1],
                          (float)w_rgb[3 * k + 2]});
    const size_t npix = rows * cols;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)npix; ++i) {
        float rv = 0.0f, gv = 0.0f, bv = 0.0f;
        for (const W &w : vw) {
            if ((size_t)i < channel_lens[w.ch]) {
                const float v = channels[w.ch][i];
                rv += v * w.rw;
                gv += v * w.gw;
                bv += v * w.bw;
            }
        }
        r[i] = rv;
        g[i] = gv;
        b[i] = bv;
    }
}

// scnr.rs:7-53
void orc_apply_scnr(float *r, float *g, float *b, size_t len, int method, float amount_in,
                    int preserve) {
    const float LUM_R = 0.2126f, LUM_G = 0.7152f, LUM_B = 0.0722f;
    const float INV_RB_WEIGHT = 1.0f / (LUM_R + LUM_B);
    const float amount = clampf(amount_in, 0.0f, 1.0f);
    if (amount < 1e-7f) return;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)len; ++i) {
        const float rv = r[i], gv = g[i], bv = b[i];
        // f32::max / f32::min ignore NaN
        const float limit = method == 0 ? (rv + bv) * 0.5f : std::fmax(rv, bv);
        const float g_corrected = std::fmin(gv, limit);
        const float g_new = gv + amount * (g_corrected - gv);
        const float delta_g = gv - g_new;
        if (preserve && delta_g > 1e-10f && rv <= 1.0f && bv <= 1.0f) {
            const float lum_lost = LUM_G * delta_g;
            const float boost = lum_lost * INV_RB_WEIGHT;
            const float rn = rv + boost, bn = bv + boost;
            r[i] = rn > 1.0f ? 1.0f : rn;
            b[i] = bn > 1.0f ? 1.0f : bn;
        }
        g[i] = g_new;
    }
}

// compose/rgb.rs:127-130, cmd/compose/color.rs:21-49
void orc_scale(const float *data, size_t len, float factor, float *out) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)len; ++i) out[i] = data[i] * factor;
}

// curves.rs:17-50
void orc_apply_levels(const float *data, size_t len, double black, double gamma, double white,
                      float *out) {
    const bool identity = std::fabs(black) < 1e-7 && std::fabs(gamma - 1.0) < 1e-7 &&
                          std::fabs(white - 1.0) < 1e-7;
    if (identity) {
        std::memcpy(out, data, len * sizeof(float));
        return;
    }
    const double range = std::fmax(white - black, 1e-15);
    const double inv_range = 1.0 / range;
    const double inv_gamma = 1.0 / clampd(gamma, 0.01, 10.0);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)len; ++i) {
        const float v = data[i];
        if (!std::isfinite(v) || v < 0.0f) {
            out[i] = 0.0f;
            continue;
        }
        const double norm = clampd(((double)v - black) * inv_range, 0.0, 1.0);
        out[i] = (float)std::pow(norm, inv_gamma);
    }
}

// curves.rs:70-95
void orc_spline_lut(const double *points_xy, size_t n_pts, float *lut) {
    std::vector<std::pair<double, double>> sorted;
    for (size_t i = 0; i < n_pts; ++i) sorted.push_back({points_xy[2 * i], points_xy[2 * i + 1]});
    std::stable_sort(sorted.begin(), sorted.end(),
                     [](const auto &a, const auto &b) { return a.first < b.first; });
    // Vec::dedup_by(|a, b| ..): removes `a` when same_bucket(a, b) with b the previous kept.
    {
        std::vector<std::pair<double, double>> d;
        for (const auto &p : sorted) {
            if (!d.empty() && std::fabs(p.first - d.back().first) < 1e-9) continue;
            d.push_back(p);
        }
        sorted.swap(d);
    }
    if (sorted.empty() || sorted[0].first > 1e-6) sorted.insert(sorted.begin(), {0.0, 0.0});
    if (sorted.empty() || sorted.back().first < 1.0 - 1e-6) sorted.push_back({1.0, 1.0});
    const size_t n = sorted.size();
    const std::vector<double> tan = fritsch_carlson_tangents(sorted);
    for (int i = 0; i < 4096; ++i) {
        const double t = (double)i / 4095.0;
        lut[i] = (float)clampd(hermite_eval(sorted, tan, n, t), 0.0, 1.0);
    }
}

// curves.rs:104-109, 186-198
void orc_apply_curve(const float *data, size_t len, const float *lut, float *out) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)len; ++i) {
        const float v = data[i];
        if (!std::isfinite(v) || v < 0.0f) {
            out[i] = 0.0f;
            continue;
        }
        const size_t idx = f2us(clampf(v, 0.0f, 1.0f) * 4095.0f);
        out[i] = lut[std::min(idx, (size_t)4095)];
    }
}

}  // extern "C"
