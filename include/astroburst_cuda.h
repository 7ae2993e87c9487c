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
#define AB_MAX_RANKS 16        /* GPUs of one box a frame-sharded stack can span */
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
    uint32_t align; /* 1: phase-correlate + shift every frame onto frame 0 first (host-buffer calls) */
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

/* mirrors SubframeWeightConfig, core/analysis/subframe.rs:24-50
 * (defaults 1.0, 0.5, 1.0, 0.3, 8.0, 0.7, 5.0, 5) */
typedef struct {
    double fwhm_weight, eccentricity_weight, snr_weight, noise_weight;
    double max_fwhm, max_eccentricity, min_snr;
    uint64_t min_stars;
} ab_subframe_config;

/* mirrors SubframeMetrics (minus the path strings), subframe.rs:9-22 */
typedef struct {
    uint64_t star_count;
    double median_fwhm, median_eccentricity, median_snr;
    double background_median, background_sigma, noise_ratio, weight;
    int32_t accepted;
} ab_subframe_metrics;

/* mirrors AffineAlignResult, core/alignment/affine.rs:93-110.
 * method: 0 affine, 1 rigid, 2 phase_correlation, 3 identity */
typedef struct {
    ab_affine transform;
    uint64_t matched_stars, inliers;
    double residual_px;
    int32_t method;
} ab_align_result;

/* mirrors StarMaskConfig, core/imaging/star_mask.rs:7-29 (defaults 2.5, 4.0, 5.0, 1.5, 30.0, false, 0.85) */
typedef struct {
    double growth_factor, softness, detection_sigma, min_fwhm, max_fwhm;
    int32_t luminance_protect;
    double luminance_ceiling;
} ab_star_mask_config;

/* mirrors StarMaskResult (minus the mask), star_mask.rs:31-36 */
typedef struct {
    uint64_t stars_masked;
    double coverage_fraction;
} ab_star_mask_info;

/* mirrors MaskedStretchConfig, core/imaging/masked_stretch.rs:7-32
 * (defaults 10, 0.25, 2.5, 4.0, true, 0.85, 0.85, 1e-5) */
typedef struct {
    uint64_t iterations;
    double target_background, mask_growth, mask_softness;
    int32_t luminance_protect;
    double luminance_ceiling, protection_amount, convergence_threshold;
} ab_masked_stretch_config;

/* mirrors MaskedStretchResult (minus the image), masked_stretch.rs:34-42 */
typedef struct {
    uint64_t iterations_run;
    double final_background;
    uint64_t stars_masked;
    double mask_coverage;
    int32_t converged;
} ab_masked_stretch_info;

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
/* cudaDeviceEnablePeerAccess from this context's device to `peer_device` (idempotent;
 * AB_ERR_UNSUPPORTED when the two GPUs are not P2P-capable).  Only needed when several GPUs are
 * driven from ONE process; across processes ab_ipc_open enables peer access itself. */
ab_status ab_ctx_enable_peer_access(ab_ctx *ctx, int peer_device);
/* CUDA IPC: ab_ipc_export writes the 64 handle bytes (cudaIpcGetMemHandle) of a cudaMalloc'd block
 * (`dev_ptr` = its base, e.g. from ab_dev_alloc); ab_ipc_open maps another process's block under
 * this context's device with lazy peer access to the exporting GPU; ab_ipc_close unmaps it.  The
 * host exchanges the handle bytes over whatever channel it has (the reference's Rust side would
 * use its own IPC; the Python mirror uses torch.distributed).  Peer mappings feed
 * ab_stack_frame_sharded_dev / ab_reshard_frames_to_bands_dev below. */
ab_status ab_ipc_export(ab_ctx *ctx, const void *dev_ptr, void *handle64_out);
ab_status ab_ipc_open(ab_ctx *ctx, const void *handle64, void **dev_ptr);
ab_status ab_ipc_close(ab_ctx *ctx, void *dev_ptr);

/* Peer group: the collectives of the sharded paths without NCCL and without host round trips.
 * Every rank exports its context's scratch block (ab_peer_export: 64 handle bytes), the host
 * gathers the handles of all ranks in rank order (world * 64 bytes) and every rank calls
 * ab_peer_connect; from then on ab_image_stats_peer_dev reduces its counters and histograms inside
 * kernels that read the peers' blocks over NVLink (barrier flags in the same blocks; a peer that
 * never arrives makes the call fail after ~2 s instead of hanging the GPU).  All ranks must issue
 * the same sequence of peer calls.  ab_peer_disconnect before any rank destroys its context. */
ab_status ab_peer_export(ab_ctx *ctx, void *handle64_out);
ab_status ab_peer_connect(ab_ctx *ctx, const void *handles64, int world, int rank);
ab_status ab_peer_disconnect(ab_ctx *ctx);

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
/* The same stack on FITS data units as they sit in the files (load_fits_image ->
 * decode_pixels, infra/fits/reader.rs:37-101, then combine.rs:94-193 with align == false): units[f] =
 * rows*cols big-endian samples of BITPIX bitpix[f] (8 / 16 / 32 / -32 / -64) with that file's
 * BSCALE / BZERO.  Bands cross PCIe in the file's sample type (half the bytes of the f32 planes for
 * BITPIX 16) and are decoded on the device behind the copy; the result equals
 * ab_fits_decode + ab_stack_sigma_clip bit for bit.  cfg->align != 0 -> AB_ERR_UNSUPPORTED. */
ab_status ab_stack_fits_units(ab_ctx *ctx, const uint8_t *const *units, const int32_t *bitpix,
                              const double *bscale, const double *bzero, size_t n_frames, size_t rows,
                              size_t cols, const ab_stack_config *cfg, float *out,
                              uint64_t *rejected_pixels);
/* Device-resident variant.  d_frames: HOST array of n_frames DEVICE pointers.  d_rejected:
 * DEVICE u64, overwritten.  npix = rows*cols (pixels are independent, so callers may pass any
 * row band: this is the multi-GPU sharding unit). */
ab_status ab_stack_sigma_clip_dev(ab_ctx *ctx, const float *const *d_frames, size_t n_frames,
                                  size_t npix, const ab_stack_config *cfg, float *d_out,
                                  uint64_t *d_rejected, void *stream);
/* NOTE on align: align = 1 reproduces the reference bit for bit, INCLUDING its sign convention --
 * phase_correlate(reference, target) peaks at MINUS the content displacement (forward DFT,
 * a * conj(b)) and shift_image_subpixel(target, dy, dx) then moves the frame AWAY from the
 * reference, which the reference's own (never-run) tests contradict.
 * align = 2 is an opt-in correction outside parity: the correlation runs with the operands
 * swapped, so every frame is shifted ONTO frame 0.  Use 2 when registration quality matters, 1
 * when the output must equal upstream's.
 * stack_images with config.align honoured (combine.rs:118-147): every frame i >= 1 is
 * phase-correlated against frame 0 (core/stacking/align.rs:84-97) and shifted with the bicubic
 * sampler (align.rs:36-57) before the combine.  offsets: NULL or 2*n_frames int32 =
 * StackResult.offsets as (dy, dx) pairs = (round(dy), round(dx)), (0,0) for frame 0.
 * ab_stack_sigma_clip with cfg->align != 0 forwards here (offsets dropped). */
ab_status ab_stack_images(ab_ctx *ctx, const float *const *frames, size_t n_frames, size_t rows,
                          size_t cols, const ab_stack_config *cfg, float *out,
                          uint64_t *rejected_pixels, int32_t *offsets);
/* ---- frame-sharded stacks over the GPUs of one box (BASELINE config 5)
 * Partition rules: rank r of `world` owns the contiguous block of frames f with
 * ab_frame_owner(f, n_frames, world) == r (f * world / n_frames; f % world when n_frames < world),
 * stored in frame order as one [n_local, rows, cols] f32 cube, and combines row band
 * [r0, r1) = ab_row_band(rows, r, world) of ALL frames (contiguous bands, sizes differ by <= 1 row).
 * cube_bases: HOST array of `world` device pointers, entry q = base of rank q's cube as addressable
 * from this context's device (its own allocation for q == rank, an ab_ipc_open mapping otherwise).
 *
 * ab_stack_frame_sharded_dev: stack_images (combine.rs:94-193, align == false) of this rank's
 * band, reading the foreign frames straight from their owners' HBM over NVLink inside the stack
 * kernel (no staging copy, no collective on the data path).  d_out_band: (r1-r0)*cols f32.
 * StackResult.rejected_pixels: this rank's count is ADDED (system-scope atomics) to each of the
 * n_targets device words d_rejected_targets[i] -- pass every rank's counter (own + peer mappings)
 * and each rank holds the global sum once all ranks' calls have completed; the caller zeroes the
 * counters beforehand and orders steps across ranks (e.g. one counter per step in flight).
 * n_targets == 0 drops the count.  Bit-identical to the single-GPU stack of all frames.
 *
 * ab_reshard_frames_to_bands_dev: the staged alternative -- copies this rank's band of every frame
 * into d_band_cube [n_frames, r1-r0, cols] (pull copies over NVLink, one 2-D copy per owner). */
void ab_row_band(size_t rows, int rank, int world, size_t *r0, size_t *r1);
int ab_frame_owner(size_t f, size_t n_frames, int world);
ab_status ab_stack_frame_sharded_dev(ab_ctx *ctx, const float *const *cube_bases, int world,
                                     int rank, size_t n_frames, size_t rows, size_t cols,
                                     const ab_stack_config *cfg, float *d_out_band,
                                     uint64_t *const *d_rejected_targets, int n_targets,
                                     void *stream);
ab_status ab_reshard_frames_to_bands_dev(ab_ctx *ctx, const float *const *cube_bases, int world,
                                         int rank, size_t n_frames, size_t rows, size_t cols,
                                         float *d_band_cube, void *stream);
/* One pixel's sample list through the same device code path (combine.rs:14-92); n <= 256.
 * Used by the parity tests for the reference's sigma_clip_combine unit tests. */
ab_status ab_sigma_clip_combine(ab_ctx *ctx, const float *values, size_t n, float sigma_low,
                                float sigma_high, uint32_t max_iter, float *mean_out,
                                uint32_t *rejected_out);

/* ---------------------------------------------------------------- FITS data unit
 * decode_pixels, infra/fits/reader.rs:37-101: big-endian BITPIX 8 / 16 / 32 / -32 / -64 -> f32,
 * `(v as f64 * bscale + bzero) as f32` unless the scaling is identity (|bscale-1|, |bzero| < 1e-15).
 * data: n_pixels * |bitpix|/8 bytes.  Any other BITPIX -> AB_ERR_UNSUPPORTED (the reference returns
 * an empty Vec).  Header parsing and file I/O stay with the caller. */
ab_status ab_fits_decode(ab_ctx *ctx, const uint8_t *data, size_t n_pixels, int bitpix, double bscale,
                         double bzero, float *out);
ab_status ab_fits_decode_dev(ab_ctx *ctx, const uint8_t *d_data, size_t n_pixels, int bitpix,
                             double bscale, double bzero, float *d_out, void *stream);
/* write_f32 / i16 / f64_slice_as_be, infra/fits/writer.rs:82-136: f32 -> big-endian BITPIX -32, 16
 * (`((val - bzero) / bscale).clamp(i16).round() as i16`) or -64. */
ab_status ab_fits_encode(ab_ctx *ctx, const float *data, size_t n_pixels, int bitpix, double bzero,
                         double bscale, uint8_t *out);
ab_status ab_fits_encode_dev(ab_ctx *ctx, const float *d_data, size_t n_pixels, int bitpix,
                             double bzero, double bscale, uint8_t *d_out, void *stream);

/* ---------------------------------------------------------------- phase correlation
 * PhaseCorrelationResult, core/alignment/phase_correlation.rs:15-20: shift_image_subpixel(target,
 * dy, dx) registers the target on the reference. */
typedef struct ab_phase_result {
    double dx, dy, confidence;
} ab_phase_result;
/* phase_correlate, phase_correlation.rs:22-99: both images cropped to the common minimum size;
 * constant / nearly empty inputs -> zeros; <= 512 x 512 one correlation, else a 512 x 512 area
 * downsample for the coarse shift and a 512 x 512 centre crop for the refinement.  Hann window,
 * power-of-two f64 FFT (cuFFT Z2Z for the reference's rustfft), normalised cross-power spectrum,
 * peak + SNR confidence + 3-point sub-pixel refinement.  Parity with the reference is to f64 FFT
 * rounding (shifts ~1e-12 px), not bit-exact. */
ab_status ab_phase_correlate(ab_ctx *ctx, const float *ref, size_t ref_rows, size_t ref_cols,
                             const float *tgt, size_t tgt_rows, size_t tgt_cols,
                             ab_phase_result *out);
/* same-size contiguous device planes; synchronizes `stream` (the result is a host struct) */
ab_status ab_phase_correlate_dev(ab_ctx *ctx, const float *d_ref, const float *d_tgt, size_t rows,
                                 size_t cols, ab_phase_result *out, void *stream);
/* area_downsample, core/alignment/downsample.rs:6-46: mean of the finite pixels of each box
 * (f64 sum in the reference's order -> bit-exact), 0 for an empty box */
ab_status ab_area_downsample(ab_ctx *ctx, const float *src, size_t in_rows, size_t in_cols,
                             size_t out_rows, size_t out_cols, float *out);
ab_status ab_area_downsample_dev(ab_ctx *ctx, const float *d_src, size_t in_rows, size_t in_cols,
                                 size_t out_rows, size_t out_cols, float *d_out, void *stream);

/* ---------------------------------------------------------------- calibration
 * core/stacking/calibration.rs.  FITS loading stays with the caller (load_fits_image).
 *
 * ab_median_combine: median_combine_row_major (:85-123) -- per pixel the upper median
 * vals[len/2] of the finite samples, 0 when none -- fused with the per-frame preprocessing of
 * create_master_dark (:160-186, frame - bias) and create_master_flat (:199-229, frame - bias, then
 * - dark * 1.0): bias / dark may be NULL, dark_ratio multiplies the dark.  n_frames == 0 ->
 * AB_ERR_INVALID ("No ... frames provided", :126-128).  n_frames <= AB_MAX_FRAMES. */
ab_status ab_median_combine(ab_ctx *ctx, const float *const *frames, size_t n_frames, size_t rows,
                            size_t cols, const float *bias, const float *dark, float dark_ratio,
                            float *out);
/* d_frames: HOST array of DEVICE pointers; npix = any row band (the multi-GPU sharding unit). */
ab_status ab_median_combine_dev(ab_ctx *ctx, const float *const *d_frames, size_t n_frames,
                                size_t npix, const float *d_bias, const float *d_dark,
                                float dark_ratio, float *d_out, void *stream);
/* master-flat normalisation, create_master_flat :231-252, in place: finite positive pixels are
 * divided by their mean (f64 sum, f32 reciprocal), every other pixel becomes 1.0; untouched when
 * no pixel is finite and positive. */
ab_status ab_normalize_flat(ab_ctx *ctx, float *data, size_t len);
ab_status ab_normalize_flat_dev(ab_ctx *ctx, float *d_data, size_t len, void *stream);
/* sigma_clipped_mean_stack, core/imaging/calibration_pipeline.rs:317-378 (the batch pipeline's
 * clip; NOT combine.rs): every sample takes part (NaN ordered last, math/median.rs:4-13); each of
 * up to max_iterations rounds (while >= 3 samples) takes the upper median and the upper-median MAD,
 * sigma = (mad as f64 * 1.4826) as f32, stops when sigma < 1e-10, keeps the samples with
 * -sigma_low < (v - median) / sigma < sigma_high (f32, strict) and stops when nothing was removed;
 * out = f32 sum of the survivors in frame order / their count (0 when none survive);
 * rejection_counts[f] = samples of frame f rejected over the whole image (n_frames u64,
 * overwritten).  n_frames <= AB_MAX_FRAMES.  BatchStackConfig defaults: 2.5, 3.0, 5. */
ab_status ab_sigma_clipped_mean_stack(ab_ctx *ctx, const float *const *frames, size_t n_frames,
                                      size_t rows, size_t cols, float sigma_low, float sigma_high,
                                      uint32_t max_iterations, float *out,
                                      uint64_t *rejection_counts);
ab_status ab_sigma_clipped_mean_stack_dev(ab_ctx *ctx, const float *const *d_frames, size_t n_frames,
                                          size_t npix, float sigma_low, float sigma_high,
                                          uint32_t max_iterations, float *d_out,
                                          uint64_t *d_rejection_counts, void *stream);
/* CalibrationConfig (:8-13); masters may be NULL.  clamp_negative = 1: calibrate_image (:47-83,
 * `if v < 0.0 { 0.0 }`); 0: the unclamped subtract_bias / subtract_dark / divide_flat (:15-45). */
typedef struct ab_calibration_config {
    const float *master_bias;
    const float *master_dark;
    const float *master_flat;
    float dark_exposure_ratio;
    uint32_t clamp_negative;
} ab_calibration_config;
ab_status ab_calibrate_image(ab_ctx *ctx, const float *raw, size_t len,
                             const ab_calibration_config *cfg, float *out);
/* the masters in cfg are DEVICE pointers here */
ab_status ab_calibrate_image_dev(ab_ctx *ctx, const float *d_raw, size_t len,
                                 const ab_calibration_config *cfg, float *d_out, void *stream);

/* ---------------------------------------------------------------- pointwise operators
 * Each op has a host-buffer form (synchronous, copies included) and a `_dev` form (device
 * pointers, enqueued on `stream`, no synchronisation).  `len` = rows*cols. */

/* apply_stf, core/imaging/stf.rs:89-102 (== make_stf_u8_fn :122-145): f32 -> u8 */
ab_status ab_apply_stf_u8(ab_ctx *ctx, const float *data, size_t len, const ab_stf_params *p,
                          const ab_image_stats *st, uint8_t *out);
ab_status ab_apply_stf_u8_dev(ab_ctx *ctx, const float *d_data, size_t len,
                              const ab_stf_params *p, const ab_image_stats *st, uint8_t *d_out,
                              void *stream);
/* apply_stf_f32 / apply_stf_inplace, stf.rs:104-120,147-155 (variant 0); the near-duplicate in
 * core/compose/rgb.rs:191-207 that divides by clip_range (variant 1) */
ab_status ab_apply_stf_f32(ab_ctx *ctx, const float *data, size_t len, const ab_stf_params *p,
                           const ab_image_stats *st, int variant, float *out);
ab_status ab_apply_stf_f32_dev(ab_ctx *ctx, const float *d_data, size_t len,
                               const ab_stf_params *p, const ab_image_stats *st, int variant,
                               float *d_out, void *stream);
/* arcsinh_stretch_with_stats, core/imaging/stretch.rs:10-46 */
ab_status ab_arcsinh_stretch(ab_ctx *ctx, const float *data, size_t len, float dmin, float dmax,
                             float factor, float gamma, float *out);
ab_status ab_arcsinh_stretch_dev(ab_ctx *ctx, const float *d_data, size_t len, float dmin,
                                 float dmax, float factor, float gamma, float *d_out,
                                 void *stream);
/* v * factor: apply_multiplier_inplace core/compose/rgb.rs:127-130, calibrate_channel
 * cmd/compose/color.rs:21-49 (out may alias data) */
ab_status ab_scale(ab_ctx *ctx, const float *data, size_t len, float factor, float *out);
ab_status ab_scale_dev(ab_ctx *ctx, const float *d_data, size_t len, float factor, float *d_out,
                       void *stream);
/* apply_levels, core/imaging/curves.rs:31-50 */
ab_status ab_apply_levels(ab_ctx *ctx, const float *data, size_t len, const ab_levels_params *p,
                          float *out);
ab_status ab_apply_levels_dev(ab_ctx *ctx, const float *d_data, size_t len,
                              const ab_levels_params *p, float *d_out, void *stream);
/* SplineLut::from_points, curves.rs:70-95 (host scalar math): points_xy = n_pts (x,y) pairs,
 * lut = AB_CURVE_LUT_SIZE f32 */
ab_status ab_spline_lut(const double *points_xy, size_t n_pts, float *lut);
/* apply_curve, curves.rs:186-198 (truncating LUT lookup, no interpolation) */
ab_status ab_apply_curve(ab_ctx *ctx, const float *data, size_t len, const float *lut, float *out);
ab_status ab_apply_curve_dev(ab_ctx *ctx, const float *d_data, size_t len, const float *d_lut,
                             float *d_out, void *stream);
/* blend_channels, core/compose/channel_blend.rs:13-70.  Weights whose channel_idx >= n_ch are
 * skipped; pixels beyond a channel's length contribute nothing; no NaN filtering. */
ab_status ab_blend_channels(ab_ctx *ctx, const float *const *channels, const size_t *channel_lens,
                            size_t n_ch, const ab_blend_weight *w, size_t n_w, size_t rows,
                            size_t cols, float *r, float *g, float *b);
ab_status ab_blend_channels_dev(ab_ctx *ctx, const float *const *d_channels,
                                const size_t *channel_lens, size_t n_ch, const ab_blend_weight *w,
                                size_t n_w, size_t rows, size_t cols, float *d_r, float *d_g,
                                float *d_b, void *stream);
/* apply_scnr_inplace, core/imaging/scnr.rs:18-53 (the caller checks equal dims, :24-26) */
ab_status ab_apply_scnr(ab_ctx *ctx, float *r, float *g, float *b, size_t len,
                        const ab_scnr_config *cfg);
ab_status ab_apply_scnr_dev(ab_ctx *ctx, float *d_r, float *d_g, float *d_b, size_t len,
                            const ab_scnr_config *cfg, void *stream);

/* The pixel chain of apply_tone_composite_cmd, cmd/processing/curves.rs:111-151, as ONE sweep:
 * apply_stf_f32 (stf.rs:104-120) per channel -> apply_levels (curves.rs:31-50; identity params are
 * skipped like :32-34) -> apply_curve (curves.rs:186-198) when the three LUTs are given (the
 * caller passes identity LUTs for channels without a curve, as curves.rs:132-136 does) ->
 * apply_scnr_inplace (scnr.rs:18-53) when apply_scnr != 0 and amount > 1e-7.  Results equal the
 * four separate calls bit for bit; outputs may alias the inputs. */
typedef struct {
    ab_stf_params stf[3];     /* r, g, b */
    ab_image_stats stats[3];  /* the stats each STF normalises with (combined stats when linked) */
    ab_levels_params levels[3];
    const float *lut[3];      /* AB_CURVE_LUT_SIZE f32 each (device pointers for _dev), or all NULL */
    int32_t apply_scnr;
    ab_scnr_config scnr;
} ab_tone_chain;
ab_status ab_tone_composite(ab_ctx *ctx, const float *r, const float *g, const float *b, size_t len,
                            const ab_tone_chain *chain, float *out_r, float *out_g, float *out_b);
ab_status ab_tone_composite_dev(ab_ctx *ctx, const float *d_r, const float *d_g, const float *d_b,
                                size_t len, const ab_tone_chain *chain, float *d_out_r,
                                float *d_out_g, float *d_out_b, void *stream);

/* ---------------------------------------------------------------- statistics / histogram / STF
 * mode: 0 compute_image_stats (core/imaging/stats.rs:15-23: exact select <= 4,000,000 px,
 *         65,536-bin 3-pass histogram algorithm above), 1 exact path (stats.rs:43-73),
 *         2 histogram path (stats.rs:75-210), 3 compute_image_stats_with_known_range
 *         (stats.rs:25-41, uses known_min/known_max).  The result struct is written on the host;
 *         the call synchronizes (the percentile walks are host scalar code, as in the reference). */
ab_status ab_compute_image_stats(ab_ctx *ctx, const float *data, size_t len, int mode,
                                 double known_min, double known_max, ab_image_stats *out);
ab_status ab_image_stats_dev(ab_ctx *ctx, const float *d_data, size_t len, int mode,
                             double known_min, double known_max, ab_image_stats *out,
                             void *stream);
/* Row-band sharded image statistics: each rank passes its band; `allreduce` reduces a
 * HOST buffer in place over all ranks (dtype 0: f64 sum, 1: u64 sum, 2: f64 min, 3: f64 max;
 * returns 0 on success) -- e.g. a torch.distributed / NCCL all-reduce.  Always the histogram
 * algorithm (stats.rs:75-210, the one the reference uses above 4,000,000 px); use_known_range
 * selects compute_image_stats_with_known_range (stats.rs:25-41).  All ranks get the same result. */
typedef int (*ab_allreduce_fn)(void *user, void *buf, size_t count, int dtype);
ab_status ab_image_stats_sharded_dev(ab_ctx *ctx, const float *d_data, size_t len,
                                     ab_allreduce_fn allreduce, void *user, int use_known_range,
                                     double known_min, double known_max, ab_image_stats *out,
                                     void *stream);
/* The same with the reductions on the DEVICE and on `stream`: `allreduce` all-reduces a device
 * buffer in place over all ranks, asynchronously on `stream` (dtype 0: f64, 1: u64, 2: u32; op 0:
 * sum, 1: min, 2: max; returns 0 on success) -- one ncclAllReduce per call.  The sweeps and the
 * device-side percentile walks then follow each other without a host round trip; the call
 * synchronises once, for the result.  Same results as ab_image_stats_sharded_dev. */
typedef int (*ab_allreduce_dev_fn)(void *user, void *d_buf, size_t count, int dtype, int op,
                                   void *stream);
ab_status ab_image_stats_sharded_chained_dev(ab_ctx *ctx, const float *d_data, size_t len,
                                             ab_allreduce_dev_fn allreduce, void *user,
                                             int use_known_range, double known_min, double known_max,
                                             ab_image_stats *out, void *stream);
/* The same over the peer group (ab_peer_connect): reductions are kernels that sum the peers'
 * counters / histograms over NVLink in rank order -- no callback, no NCCL, one synchronisation.
 * The barrier kernels spin until every rank has arrived: keep the peer collectives of ONE context
 * per GPU in flight at a time (several contexts' barriers spinning on one GPU can wait on each other
 * across ranks through streams that share a hardware queue) and enqueue them on a real stream,
 * not the NULL stream. */
ab_status ab_image_stats_peer_dev(ab_ctx *ctx, const float *d_data, size_t len, int use_known_range,
                                  double known_min, double known_max, ab_image_stats *out,
                                  void *stream);
/* build_histogram / compute_histogram_with_stats, stats.rs:373-421: n_bins must be 65536
 * (HISTOGRAM_BINS); bins = n_bins u32, edges = n_bins + 1 f64, both HOST buffers */
ab_status ab_histogram_u32(ab_ctx *ctx, const float *data, size_t len, size_t n_bins, double dmin,
                           double dmax, uint32_t *bins, double *edges);
ab_status ab_histogram_u32_dev(ab_ctx *ctx, const float *d_data, size_t len, size_t n_bins,
                               double dmin, double dmax, uint32_t *bins, double *edges,
                               void *stream);
/* downsample_histogram, stats.rs:423-444 (host) */
ab_status ab_downsample_histogram(const uint32_t *src, size_t src_len, size_t target_bins,
                                  uint32_t *out);
/* auto_stf + mtf_balance, core/imaging/stf.rs:13-47 (host scalar math) */
ab_status ab_auto_stf(const ab_image_stats *st, const ab_auto_stf_config *cfg, ab_stf_params *out);
/* linked-STF stat merge, cmd/helpers.rs:185-202 (host) */
ab_status ab_linked_stats(const ab_image_stats *r, const ab_image_stats *g, const ab_image_stats *b,
                          ab_image_stats *combined);
/* select_wb_reference, core/compose/white_balance.rs:3-20 (host): out3 = (r, g, b) factors */
ab_status ab_select_wb_reference(const ab_image_stats *r, const ab_image_stats *g,
                                 const ab_image_stats *b, double out3[3]);

/* ---------------------------------------------------------------- bicubic gathers
 * Catmull-Rom sampler of core/imaging/sampling.rs:51-80 (f64 weights, clamp-to-edge taps) under
 * the three bounds rules of its callers. */
/* warp_image, core/alignment/affine.rs:663-690: out(x,y) = sample(t(x,y)) if 0 <= s < dim-1 else 0 */
ab_status ab_warp_affine_bicubic(ab_ctx *ctx, const float *src, size_t src_rows, size_t src_cols,
                                 const ab_affine *t, size_t out_rows, size_t out_cols, float *out);
ab_status ab_warp_affine_bicubic_dev(ab_ctx *ctx, const float *d_src, size_t src_rows,
                                     size_t src_cols, const ab_affine *t, size_t out_rows,
                                     size_t out_cols, float *d_out, void *stream);
/* shift_image_subpixel, core/stacking/align.rs:36-57 (== core/alignment/pair.rs:21-39) */
ab_status ab_shift_bicubic(ab_ctx *ctx, const float *src, size_t rows, size_t cols, double dy,
                           double dx, float *out);
ab_status ab_shift_bicubic_dev(ab_ctx *ctx, const float *d_src, size_t rows, size_t cols, double dy,
                               double dx, float *d_out, void *stream);
/* resample_image, core/imaging/resample.rs:25-61; target dims of 0 -> AB_ERR_INVALID */
ab_status ab_resample_bicubic(ab_ctx *ctx, const float *src, size_t src_rows, size_t src_cols,
                              size_t target_rows, size_t target_cols, float *out);
ab_status ab_resample_bicubic_dev(ab_ctx *ctx, const float *d_src, size_t src_rows, size_t src_cols,
                                  size_t target_rows, size_t target_cols, float *d_out,
                                  void *stream);

/* ---------------------------------------------------------------- stars / alignment
 * detect_stars, core/analysis/star_detection.rs:86-258.  Up to max_stars stars (flux-descending,
 * deduplicated) are written; info->n_stars is the number found. */
ab_status ab_detect_stars(ab_ctx *ctx, const float *img, size_t rows, size_t cols,
                          double sigma_threshold, ab_star *stars, size_t max_stars,
                          ab_detection_info *info);
ab_status ab_detect_stars_dev(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols,
                              double sigma_threshold, ab_star *stars, size_t max_stars,
                              ab_detection_info *info, void *stream);
/* analyze_subframe, core/analysis/subframe.rs:52-118: detect_stars at sigma 4.0,
 * medians of fwhm / eccentricity / snr over the finite values (even count: mean of the two middle
 * ones), noise ratio, compute_weight (:120-145) and the acceptance test.  Fewer than
 * min(5, cfg->min_stars) stars -> zero metrics, weight 0, not accepted (:65-79).
 * ab_normalize_weights = normalize_weights (:147-158): divide by the largest weight when it is
 * > 1e-15 (host scalar code, no context needed). */
ab_status ab_analyze_subframe(ab_ctx *ctx, const float *img, size_t rows, size_t cols,
                              const ab_subframe_config *cfg, ab_subframe_metrics *out);
ab_status ab_analyze_subframe_dev(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols,
                                  const ab_subframe_config *cfg, ab_subframe_metrics *out,
                                  void *stream);
void ab_normalize_weights(ab_subframe_metrics *metrics, size_t n);
/* estimate_background, star_detection.rs:32-84 */
ab_status ab_estimate_background(ab_ctx *ctx, const float *img, size_t rows, size_t cols,
                                 size_t tile_size, double *median, double *sigma);
/* normalize_for_detection, core/alignment/affine.rs:24-53 */
ab_status ab_normalize_for_detection(ab_ctx *ctx, const float *img, size_t len, float *out);
/* align_channel_affine, affine.rs:129-212.  n_threads stands for rayon::current_num_threads(),
 * which seeds the reference's RANSAC streams (affine.rs:410-416).  The fallback chain ends in
 * fallback_phase_correlation (affine.rs:243-270): method = 2 with a pure translation when the
 * phase correlation is confident (>= 1.5) and within 40 % of the frame, else method = 3 (identity). */
ab_status ab_align_affine(ab_ctx *ctx, const float *ref, const float *tgt, size_t rows, size_t cols,
                          int n_threads, ab_align_result *out);
ab_status ab_align_affine_dev(ab_ctx *ctx, const float *d_ref, const float *d_tgt, size_t rows,
                              size_t cols, int n_threads, ab_align_result *out, void *stream);
/* The reference plane's star list once: align_channels_cmd (cmd/compose/blend.rs:263-296) aligns
 * every channel against the same reference, and the reference re-detects it for every pair
 * (affine.rs:133-140).  xy_out: room for 120 (x, y) pairs; *n_out = pairs written.
 * ab_align_affine_with_ref_dev(ref_xy, ...) equals ab_align_affine_dev(ref, tgt) bit for bit;
 * d_ref may be NULL when the reference plane lives on another GPU (frames spread over GPUs,
 * the star list broadcast): the phase-correlation fallback then yields identity (method 3). */
ab_status ab_align_ref_stars_dev(ab_ctx *ctx, const float *d_ref, size_t rows, size_t cols,
                                 double *xy_out, size_t *n_out, void *stream);
ab_status ab_align_affine_with_ref_dev(ab_ctx *ctx, const double *ref_xy, size_t n_ref,
                                       const float *d_ref, const float *d_tgt, size_t rows,
                                       size_t cols, int n_threads, ab_align_result *out,
                                       void *stream);
/* Both star lists given, each the output of ab_align_ref_stars_dev on its own plane (the function
 * is the aligner's view of ANY plane: normalisation, detection at 3.5 sigma, 120 brightest).  A
 * rank of a frames-across-GPUs run can therefore detect its target frames before the reference's
 * list has been broadcast.  d_ref / d_tgt feed only the phase-correlation fallback (either NULL:
 * identity, method 3).  Same result as ab_align_affine_dev(ref, tgt). */
ab_status ab_align_affine_from_stars_dev(ab_ctx *ctx, const double *ref_xy, size_t n_ref, const double *tgt_xy,
                                         size_t n_tgt, const float *d_ref, const float *d_tgt, size_t rows,
                                         size_t cols, int n_threads, ab_align_result *out, void *stream);
/* build_triangles + match_triangles, affine.rs:279-384, on (x, y) star lists (<= 120 each);
 * matches_out receives (ref x, ref y, tgt x, tgt y) per match */
ab_status ab_match_stars(ab_ctx *ctx, const double *ref_xy, size_t n_ref, const double *tgt_xy,
                         size_t n_tgt, double *matches_out, size_t max_matches, size_t *n_matches);
/* ransac_affine, affine.rs:400-517 (host scalar code); *found = 0 where the reference returns None */
ab_status ab_ransac_affine(const double *matches, size_t n, int method, int n_threads,
                           ab_align_result *out, int *found);
/* The same function with the 2,000 hypotheses of the affine model evaluated on the GPU (one thread
 * per iteration, atomicMax on (inliers, earliest iteration), refit on the device) -- the variant
 * ab_align_affine* run.  Bit-identical to ab_ransac_affine; method 1 (rigid: atan2/sin/cos) is host
 * code in both. */
ab_status ab_ransac_affine_gpu(ab_ctx *ctx, const double *matches, size_t n, int method, int n_threads,
                               ab_align_result *out, int *found);

/* ---------------------------------------------------------------- preview / export tail
 * Preview size of render_rgb_preview(_with_stf), cmd/helpers.rs:222-227: planes that fit max_dim
 * (or max_dim == 0) keep their size, else round(dim * max_dim / max(rows, cols)).max(1). */
ab_status ab_preview_size(size_t rows, size_t cols, size_t max_dim, size_t *pw, size_t *ph);
/* render_rgb, infra/render/rgb.rs:7-47 + render_rgb_preview, cmd/helpers.rs:204-262: interleaved
 * RGB8, `(v.clamp(0,1) * 255) as u8` (truncating), nearest-neighbour decimation when the planes
 * exceed max_dim.  out = ph * pw * 3 bytes (ab_preview_size).  PNG encoding stays on the host. */
ab_status ab_render_rgb_u8(ab_ctx *ctx, const float *r, const float *g, const float *b, size_t rows,
                           size_t cols, size_t max_dim, uint8_t *out);
ab_status ab_render_rgb_u8_dev(ab_ctx *ctx, const float *d_r, const float *d_g, const float *d_b,
                               size_t rows, size_t cols, size_t max_dim, uint8_t *d_out,
                               void *stream);
/* render_rgb_preview_with_stf, cmd/helpers.rs:264-322 with make_stf_u8_fn closures
 * (core/imaging/stf.rs:122-145): stf3 / stats3 = (r, g, b) parameters */
ab_status ab_render_rgb_stf_u8(ab_ctx *ctx, const float *r, const float *g, const float *b,
                               size_t rows, size_t cols, const ab_stf_params *stf3,
                               const ab_image_stats *stats3, size_t max_dim, uint8_t *out);
ab_status ab_render_rgb_stf_u8_dev(ab_ctx *ctx, const float *d_r, const float *d_g,
                                   const float *d_b, size_t rows, size_t cols,
                                   const ab_stf_params *stf3, const ab_image_stats *stats3,
                                   size_t max_dim, uint8_t *d_out, void *stream);
/* render_rgb_16bit, rgb.rs:49-98: `(v.clamp(0,1) * 65535) as u16` as big-endian bytes, interleaved;
 * out = len * 6 bytes */
ab_status ab_render_rgb_u16be(ab_ctx *ctx, const float *r, const float *g, const float *b,
                              size_t len, uint8_t *out);
ab_status ab_render_rgb_u16be_dev(ab_ctx *ctx, const float *d_r, const float *d_g,
                                  const float *d_b, size_t len, uint8_t *d_out, void *stream);
/* apply_lrgb, core/compose/lrgb.rs:4-46 (in place on r, g, b; the caller checks equal dims, :14-19) */
ab_status ab_apply_lrgb(ab_ctx *ctx, const float *l, float *r, float *g, float *b, size_t len,
                        float lightness_weight, float chrominance_weight);
ab_status ab_apply_lrgb_dev(ab_ctx *ctx, const float *d_l, float *d_r, float *d_g, float *d_b,
                            size_t len, float lightness_weight, float chrominance_weight,
                            void *stream);
/* finite_only = 0: synthesize_luminance, lrgb.rs:48-62; 1: compute_luminance,
 * core/imaging/masked_stretch.rs:134-159 (non-finite samples count as 0) */
ab_status ab_luminance(ab_ctx *ctx, const float *r, const float *g, const float *b, size_t len,
                       int finite_only, float *out);
ab_status ab_luminance_dev(ab_ctx *ctx, const float *d_r, const float *d_g, const float *d_b,
                           size_t len, int finite_only, float *d_out, void *stream);
/* generate_star_mask_from_detection, core/imaging/star_mask.rs:46-138 (stars: a detect_stars
 * result; only x, y, fwhm are read); the call synchronizes (coverage is a host scalar) */
ab_status ab_star_mask_from_stars(ab_ctx *ctx, const float *img, size_t rows, size_t cols,
                                  const ab_star *stars, size_t n_stars,
                                  const ab_star_mask_config *cfg, float *mask,
                                  ab_star_mask_info *info);
ab_status ab_star_mask_from_stars_dev(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols,
                                      const ab_star *stars, size_t n_stars,
                                      const ab_star_mask_config *cfg, float *d_mask,
                                      ab_star_mask_info *info, void *stream);
/* generate_star_mask, star_mask.rs:38-44: detect_stars(cfg->detection_sigma) then the mask */
ab_status ab_generate_star_mask(ab_ctx *ctx, const float *img, size_t rows, size_t cols,
                                const ab_star_mask_config *cfg, float *mask,
                                ab_star_mask_info *info);
ab_status ab_generate_star_mask_dev(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols,
                                    const ab_star_mask_config *cfg, float *d_mask,
                                    ab_star_mask_info *info, void *stream);
/* compute_masked_median, masked_stretch.rs:211-228 (synchronizes) */
ab_status ab_masked_median_dev(ab_ctx *ctx, const float *d_img, const float *d_mask, size_t len,
                               double *out, void *stream);
/* masked_stretch_with_mask, masked_stretch.rs:59-117; info->stars_masked / mask_coverage are the
 * caller's (they come from the StarMaskResult).  Synchronizes (the loop control is host scalar code). */
ab_status ab_masked_stretch_with_mask(ab_ctx *ctx, const float *img, size_t len, const float *mask,
                                      const ab_masked_stretch_config *cfg, float *out,
                                      ab_masked_stretch_info *info);
ab_status ab_masked_stretch_with_mask_dev(ab_ctx *ctx, const float *d_img, size_t len,
                                          const float *d_mask, const ab_masked_stretch_config *cfg,
                                          float *d_out, ab_masked_stretch_info *info, void *stream);
/* masked_stretch, masked_stretch.rs:45-57 (detect at sigma 5 -> mask -> stretch); the _dev form
 * also returns the mask in d_mask */
ab_status ab_masked_stretch(ab_ctx *ctx, const float *img, size_t rows, size_t cols,
                            const ab_masked_stretch_config *cfg, float *out,
                            ab_masked_stretch_info *info);
ab_status ab_masked_stretch_dev(ab_ctx *ctx, const float *d_img, size_t rows, size_t cols,
                                const ab_masked_stretch_config *cfg, float *d_mask, float *d_out,
                                ab_masked_stretch_info *info, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ASTROBURST_CUDA_H */
