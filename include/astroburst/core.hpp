// core.hpp -- header-only C++ mirror of the reference's `core::*` functions over the C ABI
// (include/astroburst_cuda.h).  The reference is compiled Rust and no Rust toolchain exists in the
// build image, so this is the host-language view of the boundary: same function names, argument
// meaning and error behaviour (anyhow::Error -> std::runtime_error) as the Rust functions cited.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../astroburst_cuda.h"

namespace astroburst {

struct Error : std::runtime_error {
    ab_status status;
    Error(ab_status st, const std::string &msg) : std::runtime_error(msg), status(st) {}
};

class Context {
  public:
    explicit Context(int device = -1) {
        if (ab_ctx_create(device, &ctx_) != AB_OK) throw Error(AB_ERR_CUDA, "no usable CUDA device");
    }
    ~Context() { ab_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    ab_ctx *get() const { return ctx_; }
    void check(ab_status st) const {
        if (st != AB_OK) throw Error(st, ab_last_error(ctx_));
    }

  private:
    ab_ctx *ctx_ = nullptr;
};

struct Image {  // Array2<f32>: row-major contiguous
    size_t rows = 0, cols = 0;
    std::vector<float> data;
    Image() = default;
    Image(size_t r, size_t c) : rows(r), cols(c), data(r * c) {}
};

namespace core {

namespace stacking {
// types/stacking.rs:4-28
struct StackConfig {
    float sigma_low = 3.0f, sigma_high = 3.0f;
    size_t max_iterations = 5;
    bool align = true;
};
struct StackResult {
    Image image;
    size_t frame_count = 0;
    uint64_t rejected_pixels = 0;
    std::vector<std::pair<int, int>> offsets;
};
// core/stacking/combine.rs:94-193 (frames are cropped to the common minimum like :104-113)
inline StackResult stack_images(const Context &ctx, const std::vector<Image> &images,
                                const StackConfig &config) {
    if (images.empty()) throw Error(AB_ERR_INVALID, "No images to stack");
    size_t rows = images[0].rows, cols = images[0].cols;
    for (const Image &im : images) {
        rows = im.rows < rows ? im.rows : rows;
        cols = im.cols < cols ? im.cols : cols;
    }
    std::vector<Image> cropped;
    std::vector<const float *> ptrs;
    for (const Image &im : images) {
        if (im.rows == rows && im.cols == cols) {
            ptrs.push_back(im.data.data());
        } else {
            cropped.emplace_back(rows, cols);
            for (size_t r = 0; r < rows; ++r)
                std::copy(im.data.begin() + r * im.cols, im.data.begin() + r * im.cols + cols,
                          cropped.back().data.begin() + r * cols);
        }
    }
    if (!cropped.empty()) {  // rebuild the pointer list in frame order
        ptrs.clear();
        size_t k = 0;
        for (const Image &im : images)
            ptrs.push_back(im.rows == rows && im.cols == cols ? im.data.data() : cropped[k++].data.data());
    }
    StackResult res;
    res.image = Image(rows, cols);
    res.frame_count = images.size();
    res.offsets.assign(images.size(), {0, 0});
    const ab_stack_config cfg{config.sigma_low, config.sigma_high, (uint32_t)config.max_iterations,
                              config.align ? 1u : 0u};
    ctx.check(ab_stack_sigma_clip(ctx.get(), ptrs.data(), ptrs.size(), rows, cols, &cfg,
                                  res.image.data.data(), &res.rejected_pixels));
    return res;
}

// The same stack straight from FITS data units (load_fits_image -> decode_pixels,
// infra/fits/reader.rs:37-101, then stack_images with align == false): the planes cross PCIe in
// the file's sample type and are decoded on the device (ab_stack_fits_units).
struct FitsUnit {
    const uint8_t *data = nullptr;  // rows * cols big-endian samples
    int bitpix = -32;               // 8 / 16 / 32 / -32 / -64
    double bscale = 1.0, bzero = 0.0;
};
inline StackResult stack_fits_units(const Context &ctx, const std::vector<FitsUnit> &units, size_t rows,
                                    size_t cols, const StackConfig &config) {
    if (units.empty()) throw Error(AB_ERR_INVALID, "No images to stack");
    std::vector<const uint8_t *> ptrs;
    std::vector<int32_t> bitpix;
    std::vector<double> bscale, bzero;
    for (const FitsUnit &u : units) {
        ptrs.push_back(u.data);
        bitpix.push_back(u.bitpix);
        bscale.push_back(u.bscale);
        bzero.push_back(u.bzero);
    }
    StackResult res;
    res.image = Image(rows, cols);
    res.frame_count = units.size();
    res.offsets.assign(units.size(), {0, 0});
    const ab_stack_config cfg{config.sigma_low, config.sigma_high, (uint32_t)config.max_iterations,
                              config.align ? 1u : 0u};
    ctx.check(ab_stack_fits_units(ctx.get(), ptrs.data(), bitpix.data(), bscale.data(), bzero.data(),
                                  units.size(), rows, cols, &cfg, res.image.data.data(),
                                  &res.rejected_pixels));
    return res;
}

// core/stacking/calibration.rs (frames arrive loaded; FITS I/O stays with the caller)
namespace calibration {
struct CalibrationConfig {  // :8-13
    const Image *master_bias = nullptr, *master_dark = nullptr, *master_flat = nullptr;
    float dark_exposure_ratio = 1.0f;
};
// median_combine_row_major :85-123 of (frame - bias - dark * dark_ratio)
inline Image median_combine(const Context &ctx, const std::vector<Image> &frames,
                            const Image *bias = nullptr, const Image *dark = nullptr,
                            float dark_ratio = 1.0f) {
    if (frames.empty()) throw Error(AB_ERR_INVALID, "No frames provided");
    const size_t rows = frames[0].rows, cols = frames[0].cols;
    std::vector<const float *> ptrs;
    for (const Image &f : frames) {
        if (f.rows != rows || f.cols != cols) throw Error(AB_ERR_INVALID, "Dimension mismatch");  // :140-145
        ptrs.push_back(f.data.data());
    }
    Image out(rows, cols);
    ctx.check(ab_median_combine(ctx.get(), ptrs.data(), ptrs.size(), rows, cols,
                                bias ? bias->data.data() : nullptr, dark ? dark->data.data() : nullptr,
                                dark_ratio, out.data.data()));
    return out;
}
inline Image create_master_bias(const Context &ctx, const std::vector<Image> &bias) {  // :125-152
    return median_combine(ctx, bias);
}
inline Image create_master_dark(const Context &ctx, const std::vector<Image> &dark,
                                const Image *master_bias) {  // :154-190
    return median_combine(ctx, dark, master_bias);
}
inline Image create_master_flat(const Context &ctx, const std::vector<Image> &flat,
                                const Image *master_bias, const Image *master_dark) {  // :192-256
    Image m = median_combine(ctx, flat, master_bias, master_dark, 1.0f);
    ctx.check(ab_normalize_flat(ctx.get(), m.data.data(), m.data.size()));
    return m;
}
inline Image calibrate_image(const Context &ctx, const Image &raw, const CalibrationConfig &c) {  // :47-83
    const ab_calibration_config cfg{c.master_bias ? c.master_bias->data.data() : nullptr,
                                    c.master_dark ? c.master_dark->data.data() : nullptr,
                                    c.master_flat ? c.master_flat->data.data() : nullptr,
                                    c.dark_exposure_ratio, 1u};
    Image out(raw.rows, raw.cols);
    ctx.check(ab_calibrate_image(ctx.get(), raw.data.data(), raw.data.size(), &cfg, out.data.data()));
    return out;
}
}  // namespace calibration
}  // namespace stacking

namespace imaging {
using ImageStats = ab_image_stats;  // types/image.rs:2-10
using StfParams = ab_stf_params;    // types/image.rs:35-40
// core/imaging/stats.rs:15-23
inline ImageStats compute_image_stats(const Context &ctx, const Image &img) {
    ImageStats st{};
    ctx.check(ab_compute_image_stats(ctx.get(), img.data.data(), img.data.size(), 0, 0.0, 0.0, &st));
    return st;
}
// core/imaging/stf.rs:13-39 with AutoStfConfig::default()
inline StfParams auto_stf(const ImageStats &st, double target_bg = 0.25, double shadow_k = -2.8) {
    const ab_auto_stf_config cfg{target_bg, shadow_k};
    StfParams p{};
    ab_auto_stf(&st, &cfg, &p);
    return p;
}
// core/imaging/stf.rs:89-102
inline std::vector<uint8_t> apply_stf(const Context &ctx, const Image &img, const StfParams &p,
                                      const ImageStats &st) {
    std::vector<uint8_t> out(img.data.size());
    ctx.check(ab_apply_stf_u8(ctx.get(), img.data.data(), img.data.size(), &p, &st, out.data()));
    return out;
}
// The pixel chain of apply_tone_composite_cmd (cmd/processing/curves.rs:111-151) in one sweep:
// STF -> levels -> curves (three 4096-entry LUTs or none) -> SCNR.  r, g, b are updated in place.
struct ToneChain {
    StfParams stf[3];
    ImageStats stats[3];
    ab_levels_params levels[3] = {{0.0, 1.0, 1.0}, {0.0, 1.0, 1.0}, {0.0, 1.0, 1.0}};
    const std::vector<float> *lut[3] = {nullptr, nullptr, nullptr};
    bool apply_scnr = false;
    ab_scnr_config scnr{0, 1.0f, 0};
};
inline void apply_tone_composite(const Context &ctx, Image &r, Image &g, Image &b, const ToneChain &t) {
    if (g.data.size() != r.data.size() || b.data.size() != r.data.size())
        throw Error(AB_ERR_INVALID, "channel dimension mismatch");
    ab_tone_chain c{};
    for (int k = 0; k < 3; ++k) {
        c.stf[k] = t.stf[k];
        c.stats[k] = t.stats[k];
        c.levels[k] = t.levels[k];
        c.lut[k] = t.lut[k] ? t.lut[k]->data() : nullptr;
    }
    c.apply_scnr = t.apply_scnr ? 1 : 0;
    c.scnr = t.scnr;
    ctx.check(ab_tone_composite(ctx.get(), r.data.data(), g.data.data(), b.data.data(), r.data.size(), &c,
                                r.data.data(), g.data.data(), b.data.data()));
}
}  // namespace imaging

namespace alignment {
using AffineTransform = ab_affine;  // affine.rs:56-63
// core/alignment/affine.rs:663-690
inline Image warp_image(const Context &ctx, const Image &img, const AffineTransform &t,
                        size_t out_rows, size_t out_cols) {
    Image out(out_rows, out_cols);
    ctx.check(ab_warp_affine_bicubic(ctx.get(), img.data.data(), img.rows, img.cols, &t, out_rows,
                                     out_cols, out.data.data()));
    return out;
}
// core/alignment/affine.rs:129-212
inline ab_align_result align_channel_affine(const Context &ctx, const Image &reference,
                                            const Image &target, int n_threads) {
    ab_align_result r{};
    ctx.check(ab_align_affine(ctx.get(), reference.data.data(), target.data.data(), reference.rows,
                              reference.cols, n_threads, &r));
    return r;
}
}  // namespace alignment

namespace preview {
// cmd/helpers.rs:204-262 / infra/render/rgb.rs:7-47 without the PNG encoder: interleaved RGB8
inline std::vector<uint8_t> render_rgb_preview(const Context &ctx, const Image &r, const Image &g,
                                               const Image &b, size_t max_dim, size_t *pw = nullptr,
                                               size_t *ph = nullptr) {
    size_t w = 0, h = 0;
    ab_preview_size(r.rows, r.cols, max_dim, &w, &h);
    std::vector<uint8_t> out(w * h * 3);
    ctx.check(ab_render_rgb_u8(ctx.get(), r.data.data(), g.data.data(), b.data.data(), r.rows, r.cols,
                               max_dim, out.data()));
    if (pw) *pw = w;
    if (ph) *ph = h;
    return out;
}
// core/compose/lrgb.rs:4-46
inline void apply_lrgb(const Context &ctx, const Image &l, Image &r, Image &g, Image &b,
                       float lightness_weight, float chrominance_weight) {
    if (l.rows != r.rows || l.cols != r.cols || g.data.size() != r.data.size() ||
        b.data.size() != r.data.size())
        throw Error(AB_ERR_INVALID, "L dimensions do not match RGB");
    ctx.check(ab_apply_lrgb(ctx.get(), l.data.data(), r.data.data(), g.data.data(), b.data.data(),
                            r.data.size(), lightness_weight, chrominance_weight));
}
// core/imaging/masked_stretch.rs:7-57
struct MaskedStretchConfig {
    size_t iterations = 10;
    double target_background = 0.25, mask_growth = 2.5, mask_softness = 4.0;
    bool luminance_protect = true;
    double luminance_ceiling = 0.85, protection_amount = 0.85, convergence_threshold = 1e-5;
};
struct MaskedStretchResult {
    Image image;
    ab_masked_stretch_info info{};
};
inline MaskedStretchResult masked_stretch(const Context &ctx, const Image &img,
                                          const MaskedStretchConfig &c = MaskedStretchConfig()) {
    const ab_masked_stretch_config cfg{c.iterations, c.target_background, c.mask_growth,
                                       c.mask_softness, c.luminance_protect ? 1 : 0,
                                       c.luminance_ceiling, c.protection_amount,
                                       c.convergence_threshold};
    MaskedStretchResult out;
    out.image = Image(img.rows, img.cols);
    ctx.check(ab_masked_stretch(ctx.get(), img.data.data(), img.rows, img.cols, &cfg,
                                out.image.data.data(), &out.info));
    return out;
}
}  // namespace preview

}  // namespace core
}  // namespace astroburst
