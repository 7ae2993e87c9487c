/* oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a CPU restatement (C++17, OpenMP
 * row-parallel where the reference uses rayon) of the AstroBurst per-pixel hot
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product (astroburst_b200 and
 * libastroburst_cuda.so) never does.
 *
 * PARITY PINNING: the reference is Rust and no Rust toolchain exists in the
 * build image, so the reference itself cannot be executed here.  The oracle is
 * pinned against every known-answer unit test the reference holds for this
 * path (ported in tests/test_oracle_kat.py).  The
 * reference ships NO golden output arrays, so pixel-level parity is pinned
 * only by this restatement ("parity pinned by KATs, not by golden arrays").
 *
 * Every function cites the reference file:line it follows
 * (paths relative to src-tauri/src/).
 */
#ifndef AB_ORACLE_H
#define AB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mirrors types/image.rs:2-10 */
typedef struct {
    double min, max, median, mad, sigma, mean;
    uint64_t valid_count;
} orc_image_stats;

/* mirrors types/image.rs:35-40 */
typedef struct {
    double shadow, midtone, highlight;
} orc_stf_params;

/* mirrors core/analysis/star_detection.rs:11-20 */
typedef struct {
    double x, y, flux, fwhm, eccentricity, peak;
    uint64_t npix;
    double snr;
} orc_star;

/* mirrors core/alignment/affine.rs:56-63 */
typedef struct {
    double a, b, tx, c, d, ty;
} orc_affine;

/* mirrors core/alignment/affine.rs:93-99; method: 0 affine 1 rigid 2 phase_correlation 3 identity */
typedef struct {
    orc_affine transform;
    uint64_t matched_stars, inliers;
    double residual_px;
    int32_t method;
} orc_align_result;

int orc_num_threads(void);
void orc_set_num_threads(int n);

/* ---- stacking (core/stacking/combine.rs, math/median.rs, math/sigma_clip.rs) ---- */
float orc_sigma_clip_combine(float *values, size_t n, float sigma_low, float sigma_high,
                             size_t max_iter, uint32_t *rejected);
int orc_stack_images(const float *const *frames, size_t n_frames, size_t rows, size_t cols,
                     float sigma_low, float sigma_high, size_t max_iter, float *out,
                     uint64_t *rejected_pixels);
double orc_exact_median_mut(float *data, size_t n);
float orc_median_f32_mut(float *data, size_t n);
float orc_exact_mad_mut(float *data, size_t n, float median);
/* values is consumed (retain); returns surviving length through *n */
void orc_sigma_clipped_stats(float *values, size_t *n, float kappa, size_t iterations,
                             double *median, double *sigma);

/* ---- phase correlation (core/alignment/phase_correlation.rs, downsample.rs, math) ---- */
void orc_area_downsample(const float *src, size_t in_rows, size_t in_cols, size_t out_rows,
                         size_t out_cols, float *out);
void orc_phase_correlate(const float *ref, size_t ref_rows, size_t ref_cols, const float *tgt,
                         size_t tgt_rows, size_t tgt_cols, double *dx, double *dy, double *conf);

/* ---- FITS data-unit codec (infra/fits/reader.rs:37-101, writer.rs:82-136) ---- */
int orc_fits_decode(const uint8_t *data, size_t n_pixels, int bitpix, double bscale, double bzero, float *out);
int orc_fits_encode(const float *data, size_t n_pixels, int bitpix, double bzero, double bscale, uint8_t *out);

/* ---- calibration (core/stacking/calibration.rs) ---- */
int orc_sigma_clipped_mean_stack(const float *const *frames, size_t n_frames, size_t rows, size_t cols,
                                 float sigma_low, float sigma_high, size_t max_iterations, float *out,
                                 uint64_t *rejection_counts);
int orc_median_combine(const float *const *frames, size_t n_frames, size_t rows, size_t cols,
                       const float *bias, const float *dark, float dark_ratio, float *out);
void orc_normalize_flat(float *data, size_t len);
void orc_calibrate_image(const float *raw, size_t len, const float *bias, const float *dark,
                         const float *flat, float dark_ratio, int clamp_negative, float *out);

/* ---- stats / histogram / STF params (core/imaging/stats.rs, stf.rs) ---- */
void orc_compute_image_stats(const float *data, size_t len, orc_image_stats *out);
void orc_compute_image_stats_exact(const float *data, size_t len, orc_image_stats *out);
void orc_compute_image_stats_hist(const float *data, size_t len, orc_image_stats *out);
void orc_compute_image_stats_with_known_range(const float *data, size_t len, double known_min,
                                              double known_max, orc_image_stats *out);
/* bins: n_bins u32; edges: n_bins+1 f64 */
void orc_build_histogram(const float *data, size_t len, size_t n_bins, double dmin, double dmax,
                         uint32_t *bins, double *edges);
void orc_downsample_histogram(const uint32_t *src, size_t src_len, size_t target_bins,
                              uint32_t *out);
void orc_auto_stf(const orc_image_stats *stats, double target_bg, double shadow_k,
                  orc_stf_params *out);
void orc_linked_stats(const orc_image_stats *r, const orc_image_stats *g, const orc_image_stats *b,
                      orc_image_stats *combined);
void orc_select_wb_reference(const orc_image_stats *r, const orc_image_stats *g,
                             const orc_image_stats *b, double out3[3]);

/* ---- pointwise (stf.rs, stretch.rs, scnr.rs, curves.rs, channel_blend.rs, compose/rgb.rs) ---- */
void orc_apply_stf_u8(const float *data, size_t len, const orc_stf_params *p,
                      const orc_image_stats *st, uint8_t *out);
void orc_apply_stf_f32(const float *data, size_t len, const orc_stf_params *p,
                       const orc_image_stats *st, float *out);
/* compose/rgb.rs:191-207 variant (divides by clip_range) */
void orc_apply_stf_rgbvariant_f32(const float *data, size_t len, const orc_stf_params *p,
                                  const orc_image_stats *st, float *out);
void orc_arcsinh_stretch(const float *data, size_t len, float dmin, float dmax, float factor,
                         float gamma, float *out);
/* weights: n_w entries of (channel_idx, r, g, b) */
void orc_blend_channels(const float *const *channels, const size_t *channel_lens, size_t n_ch,
                        const uint64_t *w_idx, const double *w_rgb, size_t n_w, size_t rows,
                        size_t cols, float *r, float *g, float *b);
/* method 0 average-neutral, 1 maximum-neutral */
void orc_apply_scnr(float *r, float *g, float *b, size_t len, int method, float amount,
                    int preserve_luminance);
void orc_scale(const float *data, size_t len, float factor, float *out);
void orc_apply_levels(const float *data, size_t len, double black, double gamma, double white,
                      float *out);
/* points: n_pts (x,y) pairs; lut: 4096 f32 */
void orc_spline_lut(const double *points_xy, size_t n_pts, float *lut);
void orc_apply_curve(const float *data, size_t len, const float *lut, float *out);

/* ---- resample (imaging/sampling.rs, resample.rs, alignment/affine.rs:663-690, stacking/align.rs:36-57) ---- */
double orc_catmull_rom(double t);
float orc_bicubic_sample(const float *img, size_t rows, size_t cols, double y, double x);
float orc_bilinear_sample(const float *img, size_t rows, size_t cols, double y, double x);
void orc_warp_image(const float *img, size_t src_rows, size_t src_cols, const orc_affine *t,
                    size_t out_rows, size_t out_cols, float *out);
void orc_shift_image_subpixel(const float *img, size_t rows, size_t cols, double dy, double dx,
                              float *out);
int orc_resample_image(const float *img, size_t src_rows, size_t src_cols, size_t target_rows,
                       size_t target_cols, float *out);

/* ---- stars / alignment (analysis/star_detection.rs, alignment/affine.rs) ---- */
void orc_estimate_background(const float *img, size_t rows, size_t cols, size_t tile_size,
                             double *median, double *sigma);
/* returns number of stars found (may exceed max_out; only max_out are written) */
/* SubframeWeightConfig / SubframeMetrics, core/analysis/subframe.rs:9-50 */
typedef struct {
    double fwhm_weight, eccentricity_weight, snr_weight, noise_weight, max_fwhm, max_eccentricity, min_snr;
    uint64_t min_stars;
} orc_subframe_config;
typedef struct {
    uint64_t star_count;
    double median_fwhm, median_eccentricity, median_snr, background_median, background_sigma, noise_ratio,
        weight;
    int32_t accepted;
} orc_subframe_metrics;
void orc_analyze_subframe(const float *img, size_t rows, size_t cols, const orc_subframe_config *cfg,
                          orc_subframe_metrics *out);
double orc_subframe_weight(double fwhm, double ecc, double snr, double noise, const orc_subframe_config *cfg);
void orc_normalize_weights(orc_subframe_metrics *m, size_t n);
size_t orc_detect_stars(const float *img, size_t rows, size_t cols, double sigma_threshold,
                        orc_star *out, size_t max_out, double *bg_median, double *bg_sigma);
void orc_normalize_for_detection(const float *img, size_t len, float *out);
/* n_threads emulates rayon::current_num_threads() in RANSAC seeding (affine.rs:410) */
void orc_align_channel_affine(const float *ref, const float *tgt, size_t rows, size_t cols,
                              int n_threads, orc_align_result *out);
/* pieces of the aligner, exposed for piecewise parity */
size_t orc_match_stars(const double *ref_xy, size_t n_ref, const double *tgt_xy, size_t n_tgt,
                       double *matches_out /* 4 per match */, size_t max_matches);
int orc_ransac(const double *matches, size_t n, int method, int n_threads, orc_align_result *out);
int orc_fit_affine(const double *matches, size_t n, orc_affine *out);
int orc_fit_rigid(const double *matches, size_t n, orc_affine *out);

/* ---- preview/export tail, LRGB, star mask, masked stretch ---- */
void orc_preview_size(size_t rows, size_t cols, size_t max_dim, size_t *pw, size_t *ph);
void orc_render_rgb_u8(const float *r, const float *g, const float *b, size_t rows, size_t cols,
                       size_t max_dim, uint8_t *out);
void orc_render_rgb_u16be(const float *r, const float *g, const float *b, size_t len, uint8_t *out);
void orc_render_rgb_stf_u8(const float *r, const float *g, const float *b, size_t rows, size_t cols,
                           const orc_stf_params *stf3, const orc_image_stats *st3, size_t max_dim,
                           uint8_t *out);
void orc_apply_lrgb(const float *l, float *r, float *g, float *b, size_t len, float lightness_weight,
                    float chrominance_weight);
void orc_luminance(const float *r, const float *g, const float *b, size_t len, int finite_only,
                   float *out);
size_t orc_star_mask(const float *img, size_t h, size_t w, const orc_star *stars, size_t n_stars,
                     double growth_factor, double softness, double min_fwhm, double max_fwhm,
                     int luminance_protect, double luminance_ceiling, float *mask,
                     double *coverage_fraction);
double orc_masked_median(const float *img, const float *mask, size_t len);
size_t orc_masked_stretch_with_mask(const float *img, size_t len, const float *mask,
                                    size_t iterations, double target_bg, double protection_amount,
                                    double convergence_threshold, float *out,
                                    double *final_background, int *converged);

#ifdef __cplusplus
}
#endif
#endif
