// orc_stats.cpp -- CPU oracle (TEST INFRASTRUCTURE, see oracle.h) for image
// statistics, histograms, auto-STF parameters and the white-balance scalars.
//
// Restates (paths relative to src-tauri/src/):
//   core/imaging/stats.rs:11-13     is_valid_pixel
//   core/imaging/stats.rs:15-41     compute_image_stats(+_with_known_range)
//   core/imaging/stats.rs:43-73     exact path (<= 4,000,000 px)
//   core/imaging/stats.rs:75-210    65,536-bin histogram path
//   core/imaging/stats.rs:212-353   scan/percentile helpers
//   core/imaging/stats.rs:373-444   build_histogram, downsample_histogram
//   core/imaging/stf.rs:13-47       auto_stf, mtf_balance
//   core/compose/white_balance.rs:3-20  select_wb_reference
//   cmd/helpers.rs:185-202          linked-STF stat merge
//
// The reference accumulates the f64 sum per 65,536-element chunk and reduces
// the partials in rayon's (nondeterministic) tree order (stats.rs:233-297); we
// sum chunk partials left to right.  `mean` is therefore reproducible only to
// ~1e-16 relative in the reference itself.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

constexpr float PADDING_THRESHOLD = 1e-7f;  // types/constants.rs:6
constexpr double MAD_TO_SIGMA = 1.4826;     // types/constants.rs:7
constexpr size_t HIST_BINS = 65536;         // stats.rs:8
constexpr size_t CHUNK = 65536;             // stats.rs:7

inline bool valid_px(float v) { return std::isfinite(v) && v > PADDING_THRESHOLD; }

// Rust `f64 as usize`: saturating, NaN -> 0.
inline size_t f2us(double x) {
    if (!(x > 0.0)) return 0;
    if (x >= 18446744073709551615.0) return SIZE_MAX;
    return (size_t)x;
}
inline uint64_t f2u64(double x) { return (uint64_t)f2us(x); }

inline double clampd(double x, double lo, double hi) {  // f64::clamp
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}

void stats_default(orc_image_stats *o) { std::memset(o, 0, sizeof(*o)); }

// stats.rs:212-231
void scan_minmax(const float *s, size_t len, double *mn_out, double *mx_out) {
    double mn = DBL_MAX, mx = -DBL_MAX;
#pragma omp parallel for reduction(min : mn) reduction(max : mx) schedule(static)
    for (long long i = 0; i < (long long)len; ++i) {
        const float v = s[i];
        if (valid_px(v)) {
            const double vf = (double)v;
            if (vf < mn) mn = vf;
            if (vf > mx) mx = vf;
        }
    }
    *mn_out = mn;
    *mx_out = mx;
}

// stats.rs:299-353
size_t find_percentile_bin(const std::vector<uint64_t> &hist, size_t total, double pct) {
    const uint64_t target = f2u64(std::ceil((double)total * pct));
    uint64_t cum = 0;
    for (size_t i = 0; i < hist.size(); ++i) {
        cum += hist[i];
        if (cum >= target) return i;
    }
    return hist.size() - 1;
}

double interpolate_percentile(const std::vector<uint64_t> &hist, size_t total, double pct,
                              double data_min, double bin_width) {
    const uint64_t target = f2u64(std::ceil((double)total * pct));
    uint64_t cum = 0;
    for (size_t i = 0; i < hist.size(); ++i) {
        const uint64_t count = hist[i];
        cum += count;
        if (cum >= target) {
            const uint64_t overshoot = cum - target;
            const double frac = count > 0 ? 1.0 - ((double)overshoot / (double)count) : 0.5;
            return data_min + ((double)i + frac) * bin_width;
        }
    }
    return data_min + (double)hist.size() * bin_width;
}

double resolve_rank_in_hist(const std::vector<uint64_t> &hist, uint64_t rank, double region_lo,
                            double sub_bin_width) {
    if (rank == 0) return region_lo;
    uint64_t cum = 0;
    for (size_t i = 0; i < hist.size(); ++i) {
        const uint64_t count = hist[i];
        cum += count;
        if (cum >= rank) {
            const uint64_t overshoot = cum - rank;
            const double frac = count > 0 ? 1.0 - ((double)overshoot / (double)count) : 0.5;
            return region_lo + ((double)i + frac) * sub_bin_width;
        }
    }
    return region_lo + (double)hist.size() * sub_bin_width;
}

// stats.rs:85-210
void hist_core(const float *s, size_t len, double global_min, double global_max,
               orc_image_stats *out) {
    const double range = std::fmax(global_max - global_min, 1e-30);
    const double bin_width = range / (double)HIST_BINS;
    const double inv_bin = (double)HIST_BINS / range;
    const size_t last_bin = HIST_BINS - 1;

    // pass: sum + value histogram (stats.rs:259-297)
    std::vector<uint64_t> value_hist(HIST_BINS, 0);
    double global_sum = 0.0;
    size_t total_valid = 0;
    for (size_t c0 = 0; c0 < len; c0 += CHUNK) {
        const size_t c1 = std::min(len, c0 + CHUNK);
        double sum = 0.0;
        for (size_t i = c0; i < c1; ++i) {
            const float v = s[i];
            if (valid_px(v)) {
                const double vf = (double)v;
                sum += vf;
                ++total_valid;
                const size_t idx = f2us((vf - global_min) * inv_bin);
                ++value_hist[std::min(idx, last_bin)];
            }
        }
        global_sum += sum;
    }
    if (total_valid == 0) {
        stats_default(out);
        return;
    }
    const uint64_t n = (uint64_t)total_valid;
    const double mean = global_sum / (double)n;
    const uint64_t half_count = f2u64(std::ceil((double)total_valid * 0.5));

    const size_t median_bin = find_percentile_bin(value_hist, total_valid, 0.5);
    uint64_t count_before_median = 0;
    for (size_t i = 0; i < median_bin; ++i) count_before_median += value_hist[i];
    const double median_bin_lo = global_min + (double)median_bin * bin_width;
    const double median_bin_hi = median_bin_lo + bin_width;
    const double coarse_median =
        interpolate_percentile(value_hist, total_valid, 0.5, global_min, bin_width);

    const double dev_range = range;
    const double dev_bw = dev_range / (double)HIST_BINS;
    const double dev_inv = (double)HIST_BINS / dev_range;
    const float coarse_med_f32 = (float)coarse_median;
    const double refine_range = std::fmax(median_bin_hi - median_bin_lo, 1e-30);
    const double refine_inv = (double)HIST_BINS / refine_range;

    std::vector<uint64_t> median_refine(HIST_BINS, 0), dev_hist(HIST_BINS, 0);
    for (size_t i = 0; i < len; ++i) {
        const float v = s[i];
        if (!valid_px(v)) continue;
        const double vf = (double)v;
        if (vf >= median_bin_lo && vf < median_bin_hi) {
            const size_t idx = f2us((vf - median_bin_lo) * refine_inv);
            ++median_refine[std::min(idx, last_bin)];
        }
        const float d = std::fabs(v - coarse_med_f32);  // f32 subtraction (G8)
        const size_t didx = f2us((double)d * dev_inv);
        ++dev_hist[std::min(didx, last_bin)];
    }

    const uint64_t median_rank_in_bin =
        half_count > count_before_median ? half_count - count_before_median : 0;
    const double median_refine_bw = refine_range / (double)HIST_BINS;
    const double median =
        resolve_rank_in_hist(median_refine, median_rank_in_bin, median_bin_lo, median_refine_bw);

    const size_t mad_bin = find_percentile_bin(dev_hist, total_valid, 0.5);
    const size_t expand_lo = mad_bin > 0 ? mad_bin - 1 : 0;
    const size_t expand_hi = std::min(mad_bin + 2, HIST_BINS);
    const double mad_region_lo = (double)expand_lo * dev_bw;
    const double mad_region_hi = (double)expand_hi * dev_bw;

    const float exact_med_f32 = (float)median;
    const double mad_refine_range = std::fmax(mad_region_hi - mad_region_lo, 1e-30);
    const double mad_refine_inv = (double)HIST_BINS / mad_refine_range;
    const float mad_lo_f32 = (float)mad_region_lo;
    const float mad_hi_f32 = (float)mad_region_hi;

    uint64_t count_below = 0;
    std::vector<uint64_t> mad_refine(HIST_BINS, 0);
    for (size_t i = 0; i < len; ++i) {
        const float v = s[i];
        if (!valid_px(v)) continue;
        const float dev = std::fabs(v - exact_med_f32);
        if (dev < mad_lo_f32) {
            ++count_below;
        } else if (dev < mad_hi_f32) {
            const size_t idx = f2us(((double)dev - mad_region_lo) * mad_refine_inv);
            ++mad_refine[std::min(idx, last_bin)];
        }
    }
    const uint64_t mad_rank = half_count > count_below ? half_count - count_below : 0;
    const double mad_refine_bw = mad_refine_range / (double)HIST_BINS;
    const double mad = resolve_rank_in_hist(mad_refine, mad_rank, mad_region_lo, mad_refine_bw);

    out->min = global_min;
    out->max = global_max;
    out->mean = mean;
    out->sigma = std::fmax(mad * MAD_TO_SIGMA, 1e-30);
    out->median = median;
    out->mad = mad;
    out->valid_count = n;
}

}  // namespace

extern "C" {

// stats.rs:43-73
void orc_compute_image_stats_exact(const float *s, size_t len, orc_image_stats *out) {
    double mn = DBL_MAX, mx = -DBL_MAX, gsum = 0.0;
    size_t total_valid = 0;
    for (size_t c0 = 0; c0 < len; c0 += CHUNK) {  // scan_stats, stats.rs:233-257
        const size_t c1 = std::min(len, c0 + CHUNK);
        double sum = 0.0;
        for (size_t i = c0; i < c1; ++i) {
            const float v = s[i];
            if (valid_px(v)) {
                const double vf = (double)v;
                if (vf < mn) mn = vf;
                if (vf > mx) mx = vf;
                sum += vf;
                ++total_valid;
            }
        }
        gsum += sum;
    }
    if (total_valid == 0) {
        stats_default(out);
        return;
    }
    std::vector<float> valid;
    valid.reserve(total_valid);
    for (size_t i = 0; i < len; ++i)
        if (valid_px(s[i])) valid.push_back(s[i]);
    const uint64_t n = (uint64_t)valid.size();
    const double mean = gsum / (double)n;
    const double median = orc_exact_median_mut(valid.data(), valid.size());
    const float mad_f32 = orc_exact_mad_mut(valid.data(), valid.size(), (float)median);
    const double mad = (double)mad_f32;
    out->min = mn;
    out->max = mx;
    out->mean = mean;
    out->sigma = std::fmax(mad * MAD_TO_SIGMA, 1e-30);
    out->median = median;
    out->mad = mad;
    out->valid_count = n;
}

// stats.rs:75-83
void orc_compute_image_stats_hist(const float *s, size_t len, orc_image_stats *out) {
    double mn, mx;
    scan_minmax(s, len, &mn, &mx);
    if (mn == DBL_MAX) {
        stats_default(out);
        return;
    }
    hist_core(s, len, mn, mx, out);
}

// stats.rs:15-23
void orc_compute_image_stats(const float *s, size_t len, orc_image_stats *out) {
    if (len > 4000000)
        orc_compute_image_stats_hist(s, len, out);
    else
        orc_compute_image_stats_exact(s, len, out);
}

// stats.rs:25-41
void orc_compute_image_stats_with_known_range(const float *s, size_t len, double known_min,
                                              double known_max, orc_image_stats *out) {
    if (len <= 4000000) {
        orc_compute_image_stats_exact(s, len, out);
        return;
    }
    if (!std::isfinite(known_min) || !std::isfinite(known_max) || known_min >= known_max) {
        orc_compute_image_stats_hist(s, len, out);
        return;
    }
    hist_core(s, len, known_min, known_max, out);
}

// stats.rs:378-421
void orc_build_histogram(const float *s, size_t len, size_t n_bins, double dmin, double dmax,
                         uint32_t *bins, double *edges) {
    const double range = dmax - dmin;
    std::memset(bins, 0, n_bins * sizeof(uint32_t));
    if (range < 1e-10) {
        for (size_t i = 0; i <= n_bins; ++i) edges[i] = dmin;
        return;
    }
    const double inv_bin_width = (double)n_bins / range;
    const size_t last = n_bins - 1;
    for (size_t i = 0; i < len; ++i) {
        const float v = s[i];
        if (valid_px(v)) {
            const size_t idx = f2us(((double)v - dmin) * inv_bin_width);
            ++bins[std::min(idx, last)];
        }
    }
    const double step = range / (double)n_bins;
    for (size_t i = 0; i <= n_bins; ++i) edges[i] = dmin + (double)i * step;
}

// stats.rs:423-444
void orc_downsample_histogram(const uint32_t *src, size_t src_len, size_t target_bins,
                              uint32_t *out) {
    if (target_bins >= src_len) {
        std::memcpy(out, src, src_len * sizeof(uint32_t));
        return;
    }
    const double ratio = (double)src_len / (double)target_bins;
    for (size_t i = 0; i < target_bins; ++i) {
        const size_t start = f2us((double)i * ratio);
        const size_t end = std::min(f2us((double)(i + 1) * ratio), src_len);
        uint32_t sum = 0;
        for (size_t j = start; j < end; ++j) {
            const uint64_t t = (uint64_t)sum + src[j];  // saturating_add
            sum = t > 0xffffffffull ? 0xffffffffu : (uint32_t)t;
        }
        out[i] = sum;
    }
}

// stf.rs:13-47
void orc_auto_stf(const orc_image_stats *st, double target_bg, double shadow_k,
                  orc_stf_params *out) {
    out->shadow = 0.0;
    out->midtone = 0.5;
    out->highlight = 1.0;
    if (st->valid_count == 0) return;
    const double range = std::fmax(st->max - st->min, 1e-30);
    const double median_norm = (st->median - st->min) / range;
    const double sigma_norm = st->sigma / range;
    const double shadow_norm = clampd(median_norm + shadow_k * sigma_norm, 0.0, 0.98);
    const double highlight_norm = 1.0;
    const double clip_range = std::fmax(highlight_norm - shadow_norm, 1e-15);
    const double m_clipped = clampd((median_norm - shadow_norm) / clip_range, 0.0, 1.0);
    double midtone;
    if (m_clipped <= 0.0 || m_clipped >= 1.0) {
        midtone = 0.5;
    } else {  // mtf_balance, stf.rs:41-47
        const double m = m_clipped, t = target_bg;
        const double denom = 2.0 * t * m - t - m;
        midtone = std::fabs(denom) < 1e-15 ? 0.5 : clampd(m * (t - 1.0) / denom, 0.0001, 0.9999);
    }
    out->shadow = shadow_norm;
    out->midtone = midtone;
    out->highlight = highlight_norm;
}

// cmd/helpers.rs:185-202
void orc_linked_stats(const orc_image_stats *r, const orc_image_stats *g, const orc_image_stats *b,
                      orc_image_stats *c) {
    c->min = std::fmin(std::fmin(r->min, g->min), b->min);
    c->max = std::fmax(std::fmax(r->max, g->max), b->max);
    c->mean = (r->mean + g->mean + b->mean) / 3.0;
    c->median = (r->median + g->median + b->median) / 3.0;
    c->sigma = std::sqrt((r->sigma * r->sigma + g->sigma * g->sigma + b->sigma * b->sigma) / 3.0);
    c->mad = (r->mad + g->mad + b->mad) / 3.0;
    c->valid_count = r->valid_count;
}

// core/compose/white_balance.rs:3-20
void orc_select_wb_reference(const orc_image_stats *sr, const orc_image_stats *sg,
                             const orc_image_stats *sb, double out3[3]) {
    auto stability = [](const orc_image_stats *s) {
        return s->median > 1e-10 ? s->mad / s->median : DBL_MAX;
    };
    const double stab_r = stability(sr), stab_g = stability(sg), stab_b = stability(sb);
    if (stab_r <= stab_g && stab_r <= stab_b) {
        const double m = std::fmax(sr->median, 1e-10);
        out3[0] = 1.0;
        out3[1] = m / std::fmax(sg->median, 1e-10);
        out3[2] = m / std::fmax(sb->median, 1e-10);
    } else if (stab_b <= stab_g) {
        const double m = std::fmax(sb->median, 1e-10);
        out3[0] = m / std::fmax(sr->median, 1e-10);
        out3[1] = m / std::fmax(sg->median, 1e-10);
        out3[2] = 1.0;
    } else {
        const double m = std::fmax(sg->median, 1e-10);
        out3[0] = m / std::fmax(sr->median, 1e-10);
        out3[1] = 1.0;
        out3[2] = m / std::fmax(sb->median, 1e-10);
    }
}

}  // extern "C"
