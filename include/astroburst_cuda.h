/* astroburst_cuda.h -- C ABI of libastroburst_cuda.so, the B200 (sm_100a) backend for the
 * AstroBurst per-pixel hot path.
 *
 * This is the drop-in seam one level below the Tauri commands: each entry point replaces one
 * `core::*` free function of the reference (cited per function as file:line relative to
 * src-tauri/src/).  The Rust side binds these with `extern "C"`; the
 * Tauri command / ComposeWizard surface above stays unchanged.
 *
 * Conventions
 *  - Images are row-major contiguous f32 (`Array2<f32>::as_slice()`), rows x cols.
 *  - Every function returns an ab_status; 0 is success.  Nothing throws or aborts across the
 *    boundary.  `ab_last_error(ctx)` returns the message of the last failure on that context.
 *  - Entry points without a `_dev` suffix take HOST buffers, run synchronously (copies
 *    included) and return after the result is in the caller's output buffer, matching the
 *    blocking semantics of the reference functions.  The library never retains or frees
 *    caller memory.
 *  - `_dev` entry points take DEVICE pointers and a `cudaStream_t` (as void*; NULL = the CUDA
 *    legacy default stream, as everywhere in the runtime API); they enqueue work and do NOT
 *    synchronize unless documented, so chained steps (blend -> WB -> SCNR -> tone) stay
 *    resident in HBM.
 *  - A context is bound to one CUDA device.  Entry points are re-entrant across contexts and
 *    thread-safe on one context (calls on the same context serialize).
 *  - There is no CPU fallback: without a usable CUDA device every call fails with
 *    AB_ERR_CUDA.
 */
#ifndef ASTROBURST_CUDA_H
#define ASTROBURST_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AB_ABI_VERSION 1
#define AB_MAX_FRAMES 256      /* frames per stack call (register path <= 64, deep path <= 256) */
#define AB_MAX_BLEND_WEIGHTS 32
#define AB_CURVE_LUT_SIZE 4096 /* core/imaging/curves.rs:66 */

typedef enum {
    AB_OK = 0,
    AB_ERR_INVALID = 1,     /* bad argument (the reference's bail!/Err cases) */
    AB_ERR_CUDA = 2,        /* CUDA runtime failure, message in ab_last_error */
    AB_ERR_UNSUPPORTED = 3, /* valid request outside what this build implements */
    AB_ERR_NOMEM = 4
} ab_status;

typedef struct ab_ctx ab_ctx;

/* mirrors StackConfig, types/stacking.rs:4-20 (defaults 3.0, 3.0, 5, true) */
typedef struct {
    float sigma_low;
    float sigma_high;
    uint32_t max_iterations;
    uint32_t align; /* must be 0: phase-correlation pre-alignment is SURVEY 8(f) N1 */
} ab_stack_config;

/* mirrors ImageStats, types/image.rs:2-10 */
typedef struct {
    double min, max, median, mad, sigma, mean;
    uint64_t valid_count;
} ab_image_stats;

/* mirrors StfParams, types/image.rs:35-40 */
typedef struct {
    double shadow, midtone, highlight;
} ab_stf_params;

/* mirrors AutoStfConfig, types/image.rs:57-64 (defaults 0.25, -2.8) */
typedef struct {
    double target_bg, shadow_k;
} ab_auto_stf_config;

/* mirrors ScnrConfig / ScnrMethod, types/image.rs:66-96 */
typedef struct {
    int32_t method; /* 0 AverageNeutral, 1 MaximumNeutral */
    float amount;
    int32_t preserve_luminance;
} ab_scnr_config;

/* mirrors BlendWeight, core/compose/channel_blend.rs:5-11 */
typedef struct {
    uint64_t channel_idx;
    double r_weight, g_weight, b_weight;
} ab_blend_weight;

/* mirrors LevelsParams, core/imaging/curves.rs:4-9 */
typedef struct {
    double black, gamma, white;
} ab_levels_params;

/* mirrors AffineTransform, core/alignment/affine.rs:56-63: maps OUTPUT (x,y) -> SOURCE */
typedef struct {
    double a, b, tx, c, d, ty;
} ab_affine;

/* mirrors DetectedStar, core/analysis/star_detection.rs:11-20 */
typedef struct {
    double x, y, flux, fwhm, eccentricity, peak;
    uint64_t npix;
    double snr;
} ab_star;

/* mirrors DetectionResult (minus the star Vec), star_detection.rs:23-30 */
typedef struct {
    double background_median, background_sigma, threshold_sigma;
    uint64_t image_width, image_height;
    uint64_t n_stars; /* stars found; may exceed the caller's capacity */
} ab_detection_info;

/* mirrors AffineAlignResult, core/alignment/affine.rs:93-110.
 * method: 0 affine, 1 rigid, 2 phase_correlation, 3 identity */
typedef struct {
    ab_affine transform;
    uint64_t matched_stars, inliers;
    double residual_px;
    int32_t method;
} ab_align_result;

/* ---------------------------------------------------------------- context / memory */
int ab_abi_version(void);
/* device < 0 selects the current CUDA device */
ab_status ab_ctx_create(int device, ab_ctx **out);
void ab_ctx_destroy(ab_ctx *ctx);
const char *ab_last_error(ab_ctx *ctx);
int ab_ctx_device(ab_ctx *ctx);
/* number of kernel launches issued through this context since creation */
uint64_t ab_ctx_launch_count(ab_ctx *ctx);
ab_status ab_ctx_synchronize(ab_ctx *ctx);

/* pinned host memory, so host-buffer entry points overlap copies with kernels */
ab_status ab_host_alloc(size_t bytes, void **out);
ab_status ab_host_free(void *p);
/* device-resident buffers (the `ab_buf_t` of SURVEY 8(b): cache keys map to these) */
ab_status ab_dev_alloc(ab_ctx *ctx, size_t bytes, void **out);
ab_status ab_dev_free(ab_ctx *ctx, void *p);
ab_status ab_memcpy_h2d(ab_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
ab_status ab_memcpy_d2h(ab_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);

/* ---------------------------------------------------------------- stacking
 * Replaces stack_images (core/stacking/combine.rs:94-193) with config.align == false, i.e.
 * the per-pixel gather of finite samples + sigma_clip_combine (combine.rs:14-92).
 * frames: n_frames pointers to rows*cols f32 (already cropped to the common size, as
 * combine.rs:104-113 does).  out: rows*cols f32.  rejected_pixels: StackResult.rejected_pixels.
 * n_frames == 0 -> AB_ERR_INVALID ("No images to stack", combine.rs:98-100). */
ab_status ab_stack_sigma_clip(ab_ctx *ctx, const float *const *frames, size_t n_frames,
                              size_t rows, size_t cols, const ab_stack_config *cfg, float *out,
                              uint64_t *rejected_pixels);
/* Device-resident variant.  d_frames: HOST array of n_frames DEVICE pointers.  d_rejected:
 * DEVICE u64, overwritten.  npix = rows*cols (pixels are independent, so callers may pass any
 * row band: this is the multi-GPU sharding unit). */
ab_status ab_stack_sigma_clip_dev(ab_ctx *ctx, const float *const *d_frames, size_t n_frames,
                                  size_t npix, const ab_stack_config *cfg, float *d_out,
                                  uint64_t *d_rejected, void *stream);
/* One pixel's sample list through the same device code path (combine.rs:14-92); n <= 256.
 * Used by the parity tests for the reference's sigma_clip_combine unit tests. */
ab_status ab_sigma_clip_combine(ab_ctx *ctx, const float *values, size_t n, float sigma_low,
                                float sigma_high, uint32_t max_iter, float *mean_out,
                                uint32_t *rejected_out);

#ifdef __cplusplus
}
#endif
#endif /* ASTROBURST_CUDA_H */
