// orc_stars.cpp -- CPU oracle (TEST INFRASTRUCTURE, see oracle.h) for tile
// background estimation and star detection.
//
// Restates (paths relative to src-tauri/src/):
//   core/analysis/star_detection.rs:32-84    estimate_background
//   core/analysis/star_detection.rs:86-258   detect_stars (threshold, 8-conn BFS, moments,
//                                            stable flux sort, 3-px greedy dedup)
//   core/analysis/confidence.rs:3-8          compute_detection_snr
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <deque>
#include <map>
#include <utility>
#include <vector>

#include "oracle.h"

namespace {
inline size_t f2us(double x) {  // Rust `f64 as usize`
    if (!(x > 0.0)) return 0;
    if (x >= 18446744073709551615.0) return SIZE_MAX;
    return (size_t)x;
}
inline bool f64_less(double a, double b) {  // math/median.rs:15-25, NaN last
    const bool an = std::isnan(a), bn = std::isnan(b);
    if (an || bn) return !an && bn;
    return a < b;
}
}  // namespace

extern "C" {

// star_detection.rs:32-84
void orc_estimate_background(const float *img, size_t rows, size_t cols, size_t tile_size,
                             double *median_out, double *sigma_out) {
    const size_t step = std::max(tile_size, (size_t)16);
    std::vector<std::pair<size_t, size_t>> tiles;
    for (size_t y = 0; y < rows; y += step)
        for (size_t x = 0; x < cols; x += step) tiles.push_back({y, x});

    std::vector<double> med(tiles.size()), sig(tiles.size());
    std::vector<char> ok(tiles.size(), 0);
#pragma omp parallel for schedule(dynamic, 1)
    for (long long t = 0; t < (long long)tiles.size(); ++t) {
        const size_t ty = tiles[t].first, tx = tiles[t].second;
        const size_t ye = std::min(ty + step, rows), xe = std::min(tx + step, cols);
        std::vector<float> vals;
        vals.reserve((ye - ty) * (xe - tx));
        for (size_t r = ty; r < ye; ++r)
            for (size_t c = tx; c < xe; ++c) {
                const float v = img[r * cols + c];
                if (std::isfinite(v) && v > 1e-7f) vals.push_back(v);
            }
        if (vals.size() >= 8) {
            size_t n = vals.size();
            orc_sigma_clipped_stats(vals.data(), &n, 3.0f, 2, &med[t], &sig[t]);
            ok[t] = 1;
        }
    }
    std::vector<double> medians, sigmas;
    for (size_t t = 0; t < tiles.size(); ++t)
        if (ok[t]) {
            medians.push_back(med[t]);
            sigmas.push_back(sig[t]);
        }
    if (medians.empty()) {
        *median_out = 0.0;
        *sigma_out = 1.0;
        return;
    }
    std::sort(medians.begin(), medians.end(), f64_less);
    std::sort(sigmas.begin(), sigmas.end(), f64_less);
    *median_out = medians[medians.size() / 2];
    *sigma_out = std::fmax(sigmas[sigmas.size() / 2], 1e-10);
}

// star_detection.rs:86-258
size_t orc_detect_stars(const float *img, size_t rows, size_t cols, double sigma_threshold,
                        orc_star *out, size_t max_out, double *bg_median_out,
                        double *bg_sigma_out) {
    if (rows < 3 || cols < 3) {
        *bg_median_out = 0.0;
        *bg_sigma_out = 1.0;
        return 0;
    }
    const size_t tile_size = std::min(std::max(std::min(rows, cols) / 8, (size_t)32), (size_t)256);
    double bg_median, bg_sigma;
    orc_estimate_background(img, rows, cols, tile_size, &bg_median, &bg_sigma);
    *bg_median_out = bg_median;
    *bg_sigma_out = bg_sigma;
    const double threshold = bg_median + sigma_threshold * bg_sigma;

    std::vector<char> visited(rows * cols, 0);
    std::vector<orc_star> stars;
    std::deque<std::pair<size_t, size_t>> queue;
    std::vector<std::pair<size_t, size_t>> component;
    static const int DR[8] = {-1, 1, 0, 0, -1, -1, 1, 1};
    static const int DC[8] = {0, 0, -1, 1, -1, 1, -1, 1};

    for (size_t r = 1; r + 1 < rows; ++r) {
        for (size_t c = 1; c + 1 < cols; ++c) {
            const double v = (double)img[r * cols + c];
            if (v <= threshold || visited[r * cols + c] || !std::isfinite(v)) continue;

            queue.clear();
            component.clear();
            queue.push_back({r, c});
            visited[r * cols + c] = 1;
            while (!queue.empty()) {
                const auto cur = queue.front();
                queue.pop_front();
                component.push_back(cur);
                for (int k = 0; k < 8; ++k) {
                    const long long nr = (long long)cur.first + DR[k];
                    const long long nc = (long long)cur.second + DC[k];
                    if (nr < 0 || nc < 0 || nr >= (long long)rows || nc >= (long long)cols)
                        continue;
                    const size_t ni = (size_t)nr * cols + (size_t)nc;
                    if (visited[ni]) continue;
                    const double nv = (double)img[ni];
                    if (nv > threshold && std::isfinite(nv)) {
                        visited[ni] = 1;
                        queue.push_back({(size_t)nr, (size_t)nc});
                    }
                }
            }

            const size_t npix = component.size();
            if (npix < 3 || npix > 5000) continue;
            double sum_flux = 0.0, sum_x = 0.0, sum_y = 0.0, peak_val = 0.0;
            for (const auto &p : component) {
                const double w = std::fmax((double)img[p.first * cols + p.second] - bg_median, 0.0);
                sum_flux += w;
                sum_x += (double)p.second * w;
                sum_y += (double)p.first * w;
                peak_val = std::fmax(peak_val, w);
            }
            if (sum_flux <= 0.0) continue;
            const double cx = sum_x / sum_flux, cy = sum_y / sum_flux;
            double sum_r2 = 0.0, sum_xx = 0.0, sum_yy = 0.0, sum_xy = 0.0;
            for (const auto &p : component) {
                const double w = std::fmax((double)img[p.first * cols + p.second] - bg_median, 0.0);
                const double dx = (double)p.second - cx;
                const double dy = (double)p.first - cy;
                sum_r2 += (dx * dx + dy * dy) * w;
                sum_xx += dx * dx * w;
                sum_yy += dy * dy * w;
                sum_xy += dx * dy * w;
            }
            const double sigma_star = std::sqrt(sum_r2 / (2.0 * sum_flux));
            const double fwhm = sigma_star * 2.3548200450309493;
            if (fwhm < 0.5 || fwhm > 30.0) continue;

            const double ixx = sum_xx / sum_flux, iyy = sum_yy / sum_flux, ixy = sum_xy / sum_flux;
            const double trace = ixx + iyy;
            const double det = std::fmax(ixx * iyy - ixy * ixy, 0.0);
            const double disc = std::sqrt(std::fmax((trace * trace / 4.0) - det, 0.0));
            const double lambda1 = trace / 2.0 + disc;
            const double lambda2 = std::fmax(trace / 2.0 - disc, 0.0);
            double ecc = 0.0;
            if (lambda1 > 1e-15) {
                ecc = std::sqrt(1.0 - lambda2 / lambda1);
                if (ecc < 0.0) ecc = 0.0;
                if (ecc > 1.0) ecc = 1.0;
            }
            // confidence.rs:3-8
            const double snr = bg_sigma <= DBL_EPSILON ? 0.0 : peak_val / bg_sigma;
            stars.push_back({cx, cy, sum_flux, fwhm, ecc, peak_val, (uint64_t)npix, snr});
        }
    }

    // stable sort, flux descending (partial_cmp..unwrap_or(Equal))
    std::stable_sort(stars.begin(), stars.end(),
                     [](const orc_star &a, const orc_star &b) { return a.flux > b.flux; });

    const double dedup_r2 = 9.0, cell = 3.0;
    std::map<std::pair<size_t, size_t>, std::vector<size_t>> grid;
    std::vector<orc_star> deduped;
    for (size_t i = 0; i < stars.size(); ++i) {
        const orc_star &s = stars[i];
        const size_t gx = f2us(s.x / cell), gy = f2us(s.y / cell);
        bool too_close = false;
        for (size_t ny = (gy ? gy - 1 : 0); ny <= gy + 1 && !too_close; ++ny)
            for (size_t nx = (gx ? gx - 1 : 0); nx <= gx + 1 && !too_close; ++nx) {
                auto it = grid.find({ny, nx});
                if (it == grid.end()) continue;
                for (size_t j : it->second) {
                    const double dx = s.x - stars[j].x, dy = s.y - stars[j].y;
                    if (dx * dx + dy * dy < dedup_r2) {
                        too_close = true;
                        break;
                    }
                }
            }
        if (!too_close) {
            grid[{gy, gx}].push_back(i);
            deduped.push_back(s);
        }
    }
    const size_t n = deduped.size();
    for (size_t i = 0; i < std::min(n, max_out); ++i) out[i] = deduped[i];
    return n;
}

// core/analysis/subframe.rs:52-171 -- analyze_subframe (metrics only; the path strings stay with
// the caller), compute_weight, median_of, normalize_weights.
void orc_analyze_subframe(const float *img, size_t rows, size_t cols, const orc_subframe_config *cfg,
                          orc_subframe_metrics *out) {
    std::vector<orc_star> buf(rows * cols / 3 + 16);
    double bm = 0.0, bs = 1.0;
    size_t n = orc_detect_stars(img, rows, cols, 4.0, buf.data(), buf.size(), &bm, &bs);  // :6
    n = std::min(n, buf.size());
    *out = {};
    out->star_count = n;
    out->background_median = bm;
    out->background_sigma = bs;
    if (n < std::min<size_t>(5, (size_t)cfg->min_stars)) return;  // :65-79
    auto median_of = [&](double orc_star::*field) {  // :160-171
        std::vector<double> vals;
        for (size_t i = 0; i < n; ++i)
            if (std::isfinite(buf[i].*field)) vals.push_back(buf[i].*field);
        if (vals.empty()) return 0.0;
        std::stable_sort(vals.begin(), vals.end());
        const size_t mid = vals.size() / 2;
        return vals.size() % 2 == 0 ? (vals[mid - 1] + vals[mid]) / 2.0 : vals[mid];
    };
    const double fwhm = median_of(&orc_star::fwhm), ecc = median_of(&orc_star::eccentricity);
    const double snr = median_of(&orc_star::snr);
    const double noise = bm > 1e-15 ? bs / bm : 0.0;
    out->median_fwhm = fwhm;
    out->median_eccentricity = ecc;
    out->median_snr = snr;
    out->noise_ratio = noise;
    out->weight = orc_subframe_weight(fwhm, ecc, snr, noise, cfg);
    out->accepted = (n >= (size_t)cfg->min_stars && fwhm <= cfg->max_fwhm && ecc <= cfg->max_eccentricity &&
                     snr >= cfg->min_snr)
                        ? 1
                        : 0;
}

// compute_weight, subframe.rs:120-145
double orc_subframe_weight(double fwhm, double ecc, double snr, double noise, const orc_subframe_config *c) {
    const double fwhm_score = fwhm > 0.5 ? 1.0 / fwhm : 0.0;
    const double ecc_score = 1.0 - ecc;
    const double snr_score = std::fmax(std::log(snr), 0.0);
    const double noise_score = 1.0 / (1.0 + noise * 10.0);
    const double total = c->fwhm_weight + c->eccentricity_weight + c->snr_weight + c->noise_weight;
    if (total < 1e-15) return 0.0;
    const double raw = c->fwhm_weight * fwhm_score + c->eccentricity_weight * ecc_score +
                       c->snr_weight * snr_score + c->noise_weight * noise_score;
    return std::fmax(raw / total, 0.0);
}

// normalize_weights, subframe.rs:147-158
void orc_normalize_weights(orc_subframe_metrics *m, size_t n) {
    double max_w = 0.0;
    for (size_t i = 0; i < n; ++i) max_w = std::fmax(max_w, m[i].weight);
    if (max_w > 1e-15)
        for (size_t i = 0; i < n; ++i) m[i].weight /= max_w;
}

}  // extern "C"
