// orc_align.cpp -- CPU oracle (TEST INFRASTRUCTURE, see oracle.h) for the
// star-based affine aligner.
//
// Restates core/alignment/affine.rs (paths relative to src-tauri/src/):
//   :8-22     constants              :24-53    normalize_for_detection
//   :129-212  align_channel_affine   :214-241  check_transform_sanity
//   :272-277  top_n_stars            :279-318  build_triangles
//   :320-384  match_triangles        :386-398  sort_triangle_vertices
//   :400-517  ransac_affine          :519-642  fit_affine / solve_3x3 / fit_rigid
//   :644-660  compute_residual, dist
//
// Two sources of nondeterminism in the reference are fixed here and documented:
//  * vote-tie order comes from HashMap iteration (affine.rs:351-360); the oracle
//    orders equal-vote pairs by (ref index, tgt index) ascending.
//  * the RANSAC stream depends on rayon::current_num_threads() (affine.rs:410);
//    the oracle takes that thread count as a parameter.
// The phase-correlation fallback (affine.rs:243-270) goes through orc_phasecorr.cpp (own DFT in
// place of the un-vendored rustfft).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <utility>
#include <vector>

#include <omp.h>

#include "oracle.h"

namespace {

constexpr size_t MAX_STARS = 120;
constexpr double TRIANGLE_TOLERANCE = 0.02;
constexpr size_t MIN_MATCHES_AFFINE = 6;
constexpr size_t MIN_MATCHES_RIGID = 4;
constexpr size_t RANSAC_ITERATIONS = 2000;
constexpr double RANSAC_INLIER_PX = 3.0;
constexpr double DETECTION_SIGMA = 3.5;
constexpr double MIN_TRIANGLE_SIDE = 15.0;
constexpr uint32_t MIN_VOTES = 1;
constexpr double MIN_INLIER_RATIO = 0.20;
constexpr double MAX_RESIDUAL_PX = 5.0;
constexpr double MAX_OFFSET_FRACTION = 0.40;
constexpr double MAX_ROTATION_DEG = 30.0;
constexpr double MIN_SCALE = 0.70;
constexpr double MAX_SCALE = 1.40;

struct Pt {
    double x, y;
};
struct Match {
    double rx, ry, tx, ty;
};
struct Tri {
    size_t idx[3];
    double ratio_mid, ratio_long;
};

inline double dist(Pt a, Pt b) {
    const double dx = a.x - b.x, dy = a.y - b.y;
    return std::sqrt(dx * dx + dy * dy);
}

orc_affine identity() { return {1.0, 0.0, 0.0, 0.0, 1.0, 0.0}; }

inline void amap(const orc_affine &t, double x, double y, double *px, double *py) {
    *px = t.a * x + t.b * y + t.tx;
    *py = t.c * x + t.d * y + t.ty;
}

// affine.rs:279-318 (rayon flat_map over i keeps i-major order)
std::vector<Tri> build_triangles(const std::vector<Pt> &stars) {
    std::vector<Tri> tris;
    const size_t n = stars.size();
    if (n < 3) return tris;
    const size_t limit = std::min(n, (size_t)60);
    for (size_t i = 0; i < limit; ++i)
        for (size_t j = i + 1; j < limit; ++j)
            for (size_t k = j + 1; k < limit; ++k) {
                double sides[3] = {dist(stars[i], stars[j]), dist(stars[j], stars[k]),
                                   dist(stars[i], stars[k])};
                std::stable_sort(sides, sides + 3);
                if (sides[0] < MIN_TRIANGLE_SIDE) continue;
                tris.push_back({{i, j, k}, sides[1] / sides[0], sides[2] / sides[0]});
            }
    return tris;
}

// affine.rs:386-398: vertices ordered by the length of the opposite side (stable)
void sort_triangle_vertices(const std::vector<Pt> &s, const size_t idx[3], size_t out[3]) {
    const size_t i = idx[0], j = idx[1], k = idx[2];
    std::pair<size_t, double> v[3] = {
        {i, dist(s[j], s[k])}, {j, dist(s[i], s[k])}, {k, dist(s[i], s[j])}};
    std::stable_sort(v, v + 3, [](const auto &a, const auto &b) { return a.second < b.second; });
    out[0] = v[0].first;
    out[1] = v[1].first;
    out[2] = v[2].first;
}

// affine.rs:320-384
std::vector<Match> match_triangles(const std::vector<Pt> &ref_stars,
                                   const std::vector<Pt> &tgt_stars, const std::vector<Tri> &rt,
                                   const std::vector<Tri> &tt) {
    const size_t nt = tgt_stars.size();
    std::vector<uint32_t> votes(ref_stars.size() * nt, 0);
    // pre-sorted vertex orders (pure function of the triangle)
    std::vector<size_t> rs(rt.size() * 3), ts(tt.size() * 3);
    for (size_t a = 0; a < rt.size(); ++a) sort_triangle_vertices(ref_stars, rt[a].idx, &rs[3 * a]);
    for (size_t b = 0; b < tt.size(); ++b) sort_triangle_vertices(tgt_stars, tt[b].idx, &ts[3 * b]);

    const int nth = orc_num_threads();
    std::vector<std::vector<uint32_t>> local(nth);
#pragma omp parallel num_threads(nth)
    {
        std::vector<uint32_t> &lv = local[omp_get_thread_num()];
        lv.assign(votes.size(), 0);
#pragma omp for schedule(static)
        for (long long a = 0; a < (long long)rt.size(); ++a) {
            for (size_t b = 0; b < tt.size(); ++b) {
                const double d_mid = std::fabs(rt[a].ratio_mid - tt[b].ratio_mid);
                const double d_long = std::fabs(rt[a].ratio_long - tt[b].ratio_long);
                if (d_mid > TRIANGLE_TOLERANCE || d_long > TRIANGLE_TOLERANCE) continue;
                for (int p = 0; p < 3; ++p) ++lv[rs[3 * a + p] * nt + ts[3 * b + p]];
            }
        }
    }
    for (const auto &lv : local)
        for (size_t i = 0; i < lv.size(); ++i) votes[i] += lv[i];

    struct PV {
        size_t ri, ti;
        uint32_t v;
    };
    std::vector<PV> pairs;
    for (size_t ri = 0; ri < ref_stars.size(); ++ri)
        for (size_t ti = 0; ti < nt; ++ti)
            if (votes[ri * nt + ti] > 0) pairs.push_back({ri, ti, votes[ri * nt + ti]});
    // reference: HashMap order then stable sort by votes desc; oracle: (ri,ti) asc then stable.
    std::stable_sort(pairs.begin(), pairs.end(), [](const PV &a, const PV &b) { return a.v > b.v; });

    std::vector<char> used_ref(ref_stars.size(), 0), used_tgt(nt, 0);
    std::vector<Match> matches;
    for (const PV &p : pairs) {
        if (p.v < MIN_VOTES) break;
        if (used_ref[p.ri] || used_tgt[p.ti]) continue;
        used_ref[p.ri] = 1;
        used_tgt[p.ti] = 1;
        matches.push_back({ref_stars[p.ri].x, ref_stars[p.ri].y, tgt_stars[p.ti].x,
                           tgt_stars[p.ti].y});
    }
    return matches;
}

// affine.rs:564-598
bool solve_3x3(const double a[3][3], const double b[3], double x[3]) {
    const double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) -
                       a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                       a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    if (std::fabs(det) < 1e-12) return false;
    const double inv_det = 1.0 / det;
    const double inv[3][3] = {
        {(a[1][1] * a[2][2] - a[1][2] * a[2][1]) * inv_det,
         (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * inv_det,
         (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * inv_det},
        {(a[1][2] * a[2][0] - a[1][0] * a[2][2]) * inv_det,
         (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * inv_det,
         (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * inv_det},
        {(a[1][0] * a[2][1] - a[1][1] * a[2][0]) * inv_det,
         (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * inv_det,
         (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * inv_det}};
    for (int i = 0; i < 3; ++i) x[i] = inv[i][0] * b[0] + inv[i][1] * b[1] + inv[i][2] * b[2];
    return true;
}

// affine.rs:538-562
bool solve_3x3_ls(const Match *m, size_t n, bool solve_x, double x[3]) {
    double ata[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double atb[3] = {0, 0, 0};
    for (size_t q = 0; q < n; ++q) {
        const double target = solve_x ? m[q].tx : m[q].ty;
        const double row[3] = {m[q].rx, m[q].ry, 1.0};
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) ata[i][j] += row[i] * row[j];
            atb[i] += row[i] * target;
        }
    }
    return solve_3x3(ata, atb, x);
}

// affine.rs:519-536
bool fit_affine(const Match *m, size_t n, orc_affine *t) {
    if (n < 3) return false;
    double ab[3], cd[3];
    if (!solve_3x3_ls(m, n, true, ab)) return false;
    if (!solve_3x3_ls(m, n, false, cd)) return false;
    *t = {ab[0], ab[1], ab[2], cd[0], cd[1], cd[2]};
    return true;
}

// affine.rs:600-642
bool fit_rigid(const Match *m, size_t n, orc_affine *t) {
    if (n < 2) return false;
    double rcx = 0, rcy = 0, tcx = 0, tcy = 0;
    for (size_t q = 0; q < n; ++q) {
        rcx += m[q].rx;
        rcy += m[q].ry;
        tcx += m[q].tx;
        tcy += m[q].ty;
    }
    const double nf = (double)n;
    rcx /= nf;
    rcy /= nf;
    tcx /= nf;
    tcy /= nf;
    double num = 0.0, den = 0.0;
    for (size_t q = 0; q < n; ++q) {
        const double drx = m[q].rx - rcx, dry = m[q].ry - rcy;
        const double dtx = m[q].tx - tcx, dty = m[q].ty - tcy;
        num += drx * dty - dry * dtx;
        den += drx * dtx + dry * dty;
    }
    const double theta = std::atan2(num, den);
    const double cos_t = std::cos(theta), sin_t = std::sin(theta);
    const double tx = tcx - cos_t * rcx + sin_t * rcy;
    const double ty = tcy - sin_t * rcx - cos_t * rcy;
    *t = {cos_t, -sin_t, tx, sin_t, cos_t, ty};
    return true;
}

double compute_residual(const std::vector<Match> &m, const orc_affine &t) {
    if (m.empty()) return 0.0;
    double sum = 0.0;
    for (const Match &q : m) {
        double px, py;
        amap(t, q.rx, q.ry, &px, &py);
        const double ex = px - q.tx, ey = py - q.ty;
        sum += std::sqrt(ex * ex + ey * ey);
    }
    return sum / (double)m.size();
}

// affine.rs:400-517.  method: 0 affine, 1 rigid.
bool ransac(const std::vector<Match> &matches, int method, int num_threads_in,
            orc_align_result *res) {
    const size_t n = matches.size();
    const size_t min_sample = method == 0 ? 3 : 2;
    if (n < min_sample) return false;
    const size_t num_threads = (size_t)std::max(num_threads_in, 1);
    const size_t chunk_size = (RANSAC_ITERATIONS + num_threads - 1) / num_threads;

    size_t best_inliers = 0;
    orc_affine best_t = identity();
    std::vector<char> best_mask(n, 0);

    // sequential over emulated rayon threads; reduce_with keeps the FIRST maximum
    for (size_t thread_id = 0; thread_id < num_threads; ++thread_id) {
        uint64_t state = 0xDEADBEEFCAFEBABEull + (uint64_t)thread_id * 0x9E3779B97F4A7C15ull;
        auto rnd = [&]() -> size_t {
            state ^= state << 13;
            state ^= state >> 7;
            state ^= state << 17;
            return (size_t)(state % (uint64_t)n);
        };
        size_t local_best = 0;
        orc_affine local_t = identity();
        std::vector<char> local_mask(n, 0), mask(n, 0);
        for (size_t it = 0; it < chunk_size; ++it) {
            size_t sample[3];
            size_t ns = 0;
            int attempts = 0;
            while (ns < min_sample && attempts < 20) {
                const size_t idx = rnd();
                bool has = false;
                for (size_t q = 0; q < ns; ++q) has |= sample[q] == idx;
                if (!has) sample[ns++] = idx;
                ++attempts;
            }
            if (ns < min_sample) continue;
            Match sm[3];
            for (size_t q = 0; q < ns; ++q) sm[q] = matches[sample[q]];
            orc_affine t;
            const bool ok = method == 0 ? fit_affine(sm, ns, &t) : fit_rigid(sm, ns, &t);
            if (!ok) continue;
            std::fill(mask.begin(), mask.end(), 0);
            size_t inl = 0;
            for (size_t i = 0; i < n; ++i) {
                double px, py;
                amap(t, matches[i].rx, matches[i].ry, &px, &py);
                const double ex = px - matches[i].tx, ey = py - matches[i].ty;
                const double err = std::sqrt(ex * ex + ey * ey);
                if (err < RANSAC_INLIER_PX) {
                    ++inl;
                    mask[i] = 1;
                }
            }
            if (inl > local_best) {
                local_best = inl;
                local_t = t;
                local_mask = mask;
            }
        }
        if (thread_id == 0 || local_best > best_inliers) {
            best_inliers = local_best;
            best_t = local_t;
            best_mask = local_mask;
        }
    }

    if (best_inliers < MIN_MATCHES_RIGID) return false;
    const double inlier_ratio = (double)best_inliers / (double)n;
    if (inlier_ratio < MIN_INLIER_RATIO) return false;

    std::vector<Match> inl;
    for (size_t i = 0; i < n; ++i)
        if (best_mask[i]) inl.push_back(matches[i]);
    orc_affine refined;
    const bool ok = method == 0 ? fit_affine(inl.data(), inl.size(), &refined)
                                : fit_rigid(inl.data(), inl.size(), &refined);
    if (!ok) refined = best_t;
    const double residual = compute_residual(inl, refined);
    if (residual > MAX_RESIDUAL_PX) return false;

    res->transform = refined;
    res->matched_stars = n;
    res->inliers = best_inliers;
    res->residual_px = residual;
    res->method = method;
    return true;
}

// affine.rs:214-241
bool sane(const orc_align_result &r, size_t rows, size_t cols) {
    const orc_affine &t = r.transform;
    const double max_tx = (double)cols * MAX_OFFSET_FRACTION;
    const double max_ty = (double)rows * MAX_OFFSET_FRACTION;
    if (std::fabs(t.tx) > max_tx || std::fabs(t.ty) > max_ty) return false;
    const double rot = std::fabs(std::atan2(t.c, t.a) * (180.0 / M_PI));
    if (rot > MAX_ROTATION_DEG) return false;
    const double sx = std::sqrt(t.a * t.a + t.c * t.c);
    const double sy = std::sqrt(t.b * t.b + t.d * t.d);
    if (sx < MIN_SCALE || sx > MAX_SCALE || sy < MIN_SCALE || sy > MAX_SCALE) return false;
    return true;
}

// fallback_phase_correlation, affine.rs:243-270 (orc_phasecorr.cpp stands in for rustfft)
void fallback(const float *ref, const float *tgt, size_t rows, size_t cols, orc_align_result *out) {
    out->transform = identity();
    out->matched_stars = 0;
    out->inliers = 0;
    out->residual_px = 0.0;
    out->method = 3;
    double dx, dy, conf;
    orc_phase_correlate(ref, rows, cols, tgt, rows, cols, &dx, &dy, &conf);
    const double max_tx = (double)cols * MAX_OFFSET_FRACTION, max_ty = (double)rows * MAX_OFFSET_FRACTION;
    if (std::fabs(dx) > max_tx || std::fabs(dy) > max_ty || conf < 1.5) return;
    out->transform.tx = dx;
    out->transform.ty = dy;
    out->method = 2;
}

}  // namespace

extern "C" {

// affine.rs:24-53
void orc_normalize_for_detection(const float *img, size_t len, float *out) {
    if (len == 0) return;
    const size_t step = std::max(len / 100000, (size_t)1);
    std::vector<float> samples;
    for (size_t i = 0; i < len; i += step)
        if (std::isfinite(img[i])) samples.push_back(img[i]);
    if (samples.size() < 100) {
        std::memcpy(out, img, len * sizeof(float));
        return;
    }
    std::stable_sort(samples.begin(), samples.end());
    const double lo = (double)samples[samples.size() / 100];
    const double hi = (double)samples[samples.size() * 999 / 1000];
    const double range = hi - lo;
    if (range < 1e-15) {
        std::memcpy(out, img, len * sizeof(float));
        return;
    }
    const double inv_range = 1.0 / range;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)len; ++i) {
        double v = ((double)img[i] - lo) * inv_range;
        // f64::clamp keeps NaN
        if (v < 0.0) v = 0.0;
        if (v > 1.0) v = 1.0;
        out[i] = (float)v;
    }
}

size_t orc_match_stars(const double *ref_xy, size_t n_ref, const double *tgt_xy, size_t n_tgt,
                       double *matches_out, size_t max_matches) {
    std::vector<Pt> rs(n_ref), ts(n_tgt);
    for (size_t i = 0; i < n_ref; ++i) rs[i] = {ref_xy[2 * i], ref_xy[2 * i + 1]};
    for (size_t i = 0; i < n_tgt; ++i) ts[i] = {tgt_xy[2 * i], tgt_xy[2 * i + 1]};
    const auto rt = build_triangles(rs), tt = build_triangles(ts);
    if (rt.empty() || tt.empty()) return 0;
    const auto m = match_triangles(rs, ts, rt, tt);
    for (size_t i = 0; i < std::min(m.size(), max_matches); ++i) {
        matches_out[4 * i] = m[i].rx;
        matches_out[4 * i + 1] = m[i].ry;
        matches_out[4 * i + 2] = m[i].tx;
        matches_out[4 * i + 3] = m[i].ty;
    }
    return m.size();
}

int orc_ransac(const double *matches, size_t n, int method, int n_threads, orc_align_result *out) {
    std::vector<Match> m(n);
    for (size_t i = 0; i < n; ++i)
        m[i] = {matches[4 * i], matches[4 * i + 1], matches[4 * i + 2], matches[4 * i + 3]};
    return ransac(m, method, n_threads, out) ? 1 : 0;
}

int orc_fit_affine(const double *matches, size_t n, orc_affine *out) {
    return fit_affine(reinterpret_cast<const Match *>(matches), n, out) ? 1 : 0;
}
int orc_fit_rigid(const double *matches, size_t n, orc_affine *out) {
    return fit_rigid(reinterpret_cast<const Match *>(matches), n, out) ? 1 : 0;
}

// affine.rs:129-212
void orc_align_channel_affine(const float *ref, const float *tgt, size_t rows, size_t cols,
                              int n_threads, orc_align_result *out) {
    const size_t len = rows * cols;
    std::vector<float> rn(len), tn(len);
    orc_normalize_for_detection(ref, len, rn.data());
    orc_normalize_for_detection(tgt, len, tn.data());

    std::vector<orc_star> rstars(MAX_STARS), tstars(MAX_STARS);
    double bm, bs;
    const size_t nr =
        std::min(orc_detect_stars(rn.data(), rows, cols, DETECTION_SIGMA, rstars.data(), MAX_STARS, &bm, &bs), MAX_STARS);
    const size_t nt =
        std::min(orc_detect_stars(tn.data(), rows, cols, DETECTION_SIGMA, tstars.data(), MAX_STARS, &bm, &bs), MAX_STARS);
    if (nr < MIN_MATCHES_RIGID || nt < MIN_MATCHES_RIGID) return fallback(ref, tgt, rows, cols, out);

    std::vector<Pt> rs(nr), ts(nt);
    for (size_t i = 0; i < nr; ++i) rs[i] = {rstars[i].x, rstars[i].y};
    for (size_t i = 0; i < nt; ++i) ts[i] = {tstars[i].x, tstars[i].y};
    const auto rt = build_triangles(rs), tt = build_triangles(ts);
    if (rt.empty() || tt.empty()) return fallback(ref, tgt, rows, cols, out);
    const auto matches = match_triangles(rs, ts, rt, tt);
    if (matches.size() < MIN_MATCHES_RIGID) return fallback(ref, tgt, rows, cols, out);

    orc_align_result r;
    if (matches.size() >= MIN_MATCHES_AFFINE && ransac(matches, 0, n_threads, &r) &&
        sane(r, rows, cols)) {
        *out = r;
        return;
    }
    if (matches.size() >= MIN_MATCHES_RIGID && ransac(matches, 1, n_threads, &r) &&
        sane(r, rows, cols)) {
        *out = r;
        return;
    }
    fallback(ref, tgt, rows, cols, out);
}

}  // extern "C"
