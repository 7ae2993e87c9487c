// resample.cu -- Catmull-Rom bicubic gather kernels for sm_100a.
//
// Reference semantics (paths relative to src-tauri/src/):
//   core/imaging/boundary.rs:8-20      clamp_index
//   core/imaging/sampling.rs:4-13      catmull_rom (Keys a = -0.5)
//   core/imaging/sampling.rs:51-80     bicubic_sample: f64 weights, row-then-column accumulate
//   core/alignment/affine.rs:663-690   warp_image      (0 <= s < dim-1 else 0)
//   core/stacking/align.rs:36-57       shift_image_subpixel (-0.5 <= s <= dim-0.5 else 0; clone
//                                      when |d| < 1e-12)
//   core/imaging/resample.rs:25-61     resample_image  (never skips, clamped taps)
//
// One thread per output pixel, 2-D blocks of 32 x 8 so a warp walks one output row: the 16 taps
// of neighbouring lanes overlap and are served by L1/L2 (a source row band of an output tile is
// read from HBM once).  Coordinates and weights are f64 exactly as in the reference
// (-fmad=false keeps a*x + b*y + tx as two rounded products and two rounded sums).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "ab_internal.h"

namespace {

__device__ __forceinline__ double catmull_rom(double t) {
    const double a = fabs(t);
    if (a <= 1.0) return a * a * (1.5 * a - 2.5) + 1.0;
    if (a <= 2.0) return a * (a * (2.5 - 0.5 * a) - 4.0) + 2.0;
    return 0.0;
}

__device__ __forceinline__ size_t clamp_index(long long idx, size_t len) {
    if (idx < 0) return 0;
    if (idx >= (long long)len) return len - 1;
    return (size_t)idx;
}

// bicubic_sample exactly as written in the reference (any coordinate, any plane size)
__device__ __noinline__ float bicubic_sample_generic(const float *__restrict__ s, size_t rows,
                                                     size_t cols, double y, double x) {
    const double fxf = floor(x), fyf = floor(y);
    // `as i64` saturates; NaN -> 0
    const long long ix = (x == x) ? (long long)fxf : 0;
    const long long iy = (y == y) ? (long long)fyf : 0;
    const double fx = x - (double)ix;
    const double fy = y - (double)iy;
    const double wx0 = catmull_rom(fx + 1.0), wx1 = catmull_rom(fx), wx2 = catmull_rom(fx - 1.0),
                 wx3 = catmull_rom(fx - 2.0);
    const size_t c0 = clamp_index(ix - 1, cols), c1 = clamp_index(ix, cols),
                 c2 = clamp_index(ix + 1, cols), c3 = clamp_index(ix + 2, cols);
    double val = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const size_t r = clamp_index(iy + j - 1, rows);
        const float *row = s + r * cols;
        double row_val = 0.0;
        row_val += (double)__ldg(row + c0) * wx0;
        row_val += (double)__ldg(row + c1) * wx1;
        row_val += (double)__ldg(row + c2) * wx2;
        row_val += (double)__ldg(row + c3) * wx3;
        val += row_val * catmull_rom(fy - (double)(j - 1));
    }
    return (float)val;
}

// The two polynomial pieces of catmull_rom.  For a fractional offset f in [0, 1] the four taps sit
// at |t| = f + 1, f, 1 - f, 2 - f, so the piece each tap takes is known up front; where an argument
// rounds onto the joint |t| = 1 both pieces give exactly +0, so picking the piece statically
// returns the same bits as the reference's comparisons (the f64 operations are the same ones).
__device__ __forceinline__ double cr_near(double a) { return a * a * (1.5 * a - 2.5) + 1.0; }       // |t| <= 1
__device__ __forceinline__ double cr_far(double a) { return a * (a * (2.5 - 0.5 * a) - 4.0) + 2.0; }  // 1 <= |t| <= 2

// The same value on the common path: coordinates inside a plane of < 2^31 pixels per side, so the
// tap indices are 32-bit, the fractional parts are x - floor(x) (== x - (double)(i64)floor(x)) and
// interior pixels need no clamping.  Everything else goes to bicubic_sample_generic.
// kInBounds: the caller has already established -1 <= x <= cols and -1 <= y <= rows (the warp's and
// the shift's own bounds rules), so floor() and the fractions need no range test (8 f64 compares).
template <bool kInBounds>
__device__ __forceinline__ float bicubic_sample(const float *__restrict__ s, size_t rows,
                                                size_t cols, double y, double x) {
    const double fxf = floor(x), fyf = floor(y);
    const double fx = x - fxf, fy = y - fyf;
    bool plain = rows < 0x7ffffff0u && cols < 0x7ffffff0u;
    if (!kInBounds)
        plain = plain && fxf >= -2.0 && fyf >= -2.0 && fxf <= (double)cols && fyf <= (double)rows &&
                fx >= 0.0 && fx <= 1.0 && fy >= 0.0 && fy <= 1.0;
    if (!plain) return bicubic_sample_generic(s, rows, cols, y, x);
    const int ix = (int)fxf, iy = (int)fyf, nc = (int)cols, nr = (int)rows;
    const double wx0 = cr_far(fx + 1.0), wx1 = cr_near(fx), wx2 = cr_near(fabs(fx - 1.0)),
                 wx3 = cr_far(fabs(fx - 2.0));
    const double wy[4] = {cr_far(fy + 1.0), cr_near(fy), cr_near(fabs(fy - 1.0)), cr_far(fabs(fy - 2.0))};
    double val = 0.0;
    if (ix >= 1 && iy >= 1 && ix + 2 < nc && iy + 2 < nr) {
        const float *p = s + (size_t)(iy - 1) * cols + (size_t)(ix - 1);
#pragma unroll
        for (int j = 0; j < 4; ++j, p += cols) {
            double row_val = 0.0;
            row_val += (double)__ldg(p) * wx0;
            row_val += (double)__ldg(p + 1) * wx1;
            row_val += (double)__ldg(p + 2) * wx2;
            row_val += (double)__ldg(p + 3) * wx3;
            val += row_val * wy[j];
        }
    } else {
        const int c0 = min(max(ix - 1, 0), nc - 1), c1 = min(max(ix, 0), nc - 1), c2 = min(max(ix + 1, 0), nc - 1),
                  c3 = min(max(ix + 2, 0), nc - 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float *row = s + (size_t)min(max(iy + j - 1, 0), nr - 1) * cols;
            double row_val = 0.0;
            row_val += (double)__ldg(row + c0) * wx0;
            row_val += (double)__ldg(row + c1) * wx1;
            row_val += (double)__ldg(row + c2) * wx2;
            row_val += (double)__ldg(row + c3) * wx3;
            val += row_val * wy[j];
        }
    }
    return (float)val;
}

enum { MODE_WARP = 0, MODE_SHIFT = 1, MODE_RESAMPLE = 2 };

struct GatherParams {
    double a, b, tx, c, d, ty;        // warp
    double dy, dx;                    // shift
    double scale_y, scale_x, hy, hx;  // resample
};

template <int MODE>
__global__ void __launch_bounds__(256)
gather_kernel(const float *__restrict__ src, size_t src_rows, size_t src_cols,
              float *__restrict__ out, size_t out_rows, size_t out_cols, GatherParams gp) {
    const size_t x = (size_t)blockIdx.x * 32 + threadIdx.x;
    if (x >= out_cols) return;
    // each block walks a contiguous band of output rows, 8 rows per step: consecutive steps reuse
    // the overlapping source rows from L1, and block start-up is paid once per band
    const size_t band = (((out_rows + gridDim.y - 1) / gridDim.y) + 7) / 8 * 8;
    const size_t y_begin = (size_t)blockIdx.y * band;
    const size_t y_end = y_begin + band < out_rows ? y_begin + band : out_rows;
    for (size_t y = y_begin + threadIdx.y; y < y_end; y += 8) {
        float r = 0.0f;
        if (MODE == MODE_WARP) {
            const double sx = gp.a * (double)x + gp.b * (double)y + gp.tx;
            const double sy = gp.c * (double)x + gp.d * (double)y + gp.ty;
            if (sx >= 0.0 && sy >= 0.0 && sx < (double)(src_cols - 1) && sy < (double)(src_rows - 1))
                r = bicubic_sample<true>(src, src_rows, src_cols, sy, sx);
        } else if (MODE == MODE_SHIFT) {
            const double sy = (double)y + gp.dy;
            const double sx = (double)x + gp.dx;
            // (a NaN shift passes the reference's bounds rule -- every comparison is false -- and must
            // reach the literal transcription: its NaN weights come out as 0, not NaN)
            if (!(sy < -0.5 || sy > (double)src_rows - 0.5 || sx < -0.5 ||
                  sx > (double)src_cols - 0.5))
                r = (gp.dy == gp.dy && gp.dx == gp.dx) ? bicubic_sample<true>(src, src_rows, src_cols, sy, sx)
                                                       : bicubic_sample<false>(src, src_rows, src_cols, sy, sx);
        } else {
            const double sy = (double)y * gp.scale_y + gp.hy;
            const double sx = (double)x * gp.scale_x + gp.hx;
            r = bicubic_sample<false>(src, src_rows, src_cols, sy, sx);
        }
        __stcs(out + y * out_cols + x, r);
    }
}

template <int MODE>
ab_status launch_gather(ab_ctx *ctx, const float *src, size_t sr, size_t sc, float *out, size_t orows,
                        size_t ocols, const GatherParams &gp, cudaStream_t s) {
    if (orows == 0 || ocols == 0) return AB_OK;
    dim3 block(32, 8);
    const size_t gx = (ocols + 31) / 32;
    size_t gy = (orows + 7) / 8;
    const size_t cap = std::max<size_t>(1, (size_t)ctx->sm_count * 32 / gx);  // ~8 waves of 4 blocks per SM
    if (gy > cap) gy = cap;
    if (gy > 65535) gy = 65535;
    dim3 grid((unsigned)gx, (unsigned)gy);
    gather_kernel<MODE><<<grid, block, 0, s>>>(src, sr, sc, out, orows, ocols, gp);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

// host-buffer wrapper: src plane in, out plane back (scratch slot 0)
template <typename Body>
ab_status with_src_out(ab_ctx *ctx, const float *src, size_t src_px, float *out, size_t out_px,
                       Body body) {
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    const size_t a = (src_px * 4 + 255) / 256 * 256, b = (out_px * 4 + 255) / 256 * 256;
    ab_status st = ab_scratch_reserve(ctx, 0, a + b);
    if (st != AB_OK) return st;
    float *d_src = static_cast<float *>(ctx->d_scratch[0]);
    float *d_out = reinterpret_cast<float *>(static_cast<char *>(ctx->d_scratch[0]) + a);
    AB_CUDA(ctx, cudaMemcpyAsync(d_src, src, src_px * 4, cudaMemcpyHostToDevice, ctx->stream));
    st = body(d_src, d_out, ctx->stream);
    if (st != AB_OK) return st;
    AB_CUDA(ctx, cudaMemcpyAsync(out, d_out, out_px * 4, cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}

}  // namespace

extern "C" {

ab_status ab_warp_affine_bicubic_dev(ab_ctx *ctx, const float *d_src, size_t src_rows,
                                     size_t src_cols, const ab_affine *t, size_t out_rows,
                                     size_t out_cols, float *d_out, void *stream) {
    if (!ctx || !t || !d_src || !d_out) return AB_ERR_INVALID;
    if (src_rows == 0 || src_cols == 0) return ab_fail(ctx, AB_ERR_INVALID, "empty source image");
    DeviceGuard g(ctx->device);
    GatherParams gp{};
    gp.a = t->a; gp.b = t->b; gp.tx = t->tx; gp.c = t->c; gp.d = t->d; gp.ty = t->ty;
    return launch_gather<MODE_WARP>(ctx, d_src, src_rows, src_cols, d_out, out_rows, out_cols, gp,
                                    (cudaStream_t)stream);
}
ab_status ab_warp_affine_bicubic(ab_ctx *ctx, const float *src, size_t src_rows, size_t src_cols,
                                 const ab_affine *t, size_t out_rows, size_t out_cols, float *out) {
    if (!ctx || !t || !src || !out) return AB_ERR_INVALID;
    if (src_rows == 0 || src_cols == 0) return ab_fail(ctx, AB_ERR_INVALID, "empty source image");
    GatherParams gp{};
    gp.a = t->a; gp.b = t->b; gp.tx = t->tx; gp.c = t->c; gp.d = t->d; gp.ty = t->ty;
    return with_src_out(ctx, src, src_rows * src_cols, out, out_rows * out_cols,
                        [&](const float *ds, float *dout, cudaStream_t s) {
                            return launch_gather<MODE_WARP>(ctx, ds, src_rows, src_cols, dout,
                                                            out_rows, out_cols, gp, s);
                        });
}

ab_status ab_shift_bicubic_dev(ab_ctx *ctx, const float *d_src, size_t rows, size_t cols, double dy,
                               double dx, float *d_out, void *stream) {
    if (!ctx || (rows * cols && (!d_src || !d_out))) return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (std::fabs(dy) < 1e-12 && std::fabs(dx) < 1e-12) {  // align.rs:37-39: clone
        AB_CUDA(ctx, cudaMemcpyAsync(d_out, d_src, rows * cols * 4, cudaMemcpyDeviceToDevice, s));
        return AB_OK;
    }
    GatherParams gp{};
    gp.dy = dy; gp.dx = dx;
    return launch_gather<MODE_SHIFT>(ctx, d_src, rows, cols, d_out, rows, cols, gp, s);
}
ab_status ab_shift_bicubic(ab_ctx *ctx, const float *src, size_t rows, size_t cols, double dy,
                           double dx, float *out) {
    if (!ctx || (rows * cols && (!src || !out))) return AB_ERR_INVALID;
    if (std::fabs(dy) < 1e-12 && std::fabs(dx) < 1e-12) {
        std::memcpy(out, src, rows * cols * 4);
        return AB_OK;
    }
    if (rows * cols == 0) return AB_OK;
    GatherParams gp{};
    gp.dy = dy; gp.dx = dx;
    return with_src_out(ctx, src, rows * cols, out, rows * cols,
                        [&](const float *ds, float *dout, cudaStream_t s) {
                            return launch_gather<MODE_SHIFT>(ctx, ds, rows, cols, dout, rows, cols,
                                                             gp, s);
                        });
}

static ab_status resample_setup(ab_ctx *ctx, size_t sr, size_t sc, size_t tr, size_t tc,
                                GatherParams *gp) {
    if (tr == 0 || tc == 0)
        return ab_fail(ctx, AB_ERR_INVALID, "Target dimensions must be > 0");  // resample.rs:32-34
    if (sr == 0 || sc == 0) return ab_fail(ctx, AB_ERR_INVALID, "empty source image");
    gp->scale_y = (double)sr / (double)tr;
    gp->scale_x = (double)sc / (double)tc;
    gp->hy = (gp->scale_y - 1.0) * 0.5;
    gp->hx = (gp->scale_x - 1.0) * 0.5;
    return AB_OK;
}
ab_status ab_resample_bicubic_dev(ab_ctx *ctx, const float *d_src, size_t src_rows, size_t src_cols,
                                  size_t target_rows, size_t target_cols, float *d_out,
                                  void *stream) {
    if (!ctx || !d_src || !d_out) return AB_ERR_INVALID;
    GatherParams gp{};
    ab_status st = resample_setup(ctx, src_rows, src_cols, target_rows, target_cols, &gp);
    if (st != AB_OK) return st;
    DeviceGuard g(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (target_rows == src_rows && target_cols == src_cols) {  // resample.rs:36-38: clone
        AB_CUDA(ctx, cudaMemcpyAsync(d_out, d_src, src_rows * src_cols * 4,
                                     cudaMemcpyDeviceToDevice, s));
        return AB_OK;
    }
    return launch_gather<MODE_RESAMPLE>(ctx, d_src, src_rows, src_cols, d_out, target_rows,
                                        target_cols, gp, s);
}
ab_status ab_resample_bicubic(ab_ctx *ctx, const float *src, size_t src_rows, size_t src_cols,
                              size_t target_rows, size_t target_cols, float *out) {
    if (!ctx || !src || !out) return AB_ERR_INVALID;
    GatherParams gp{};
    ab_status st = resample_setup(ctx, src_rows, src_cols, target_rows, target_cols, &gp);
    if (st != AB_OK) return st;
    if (target_rows == src_rows && target_cols == src_cols) {
        std::memcpy(out, src, src_rows * src_cols * 4);
        return AB_OK;
    }
    return with_src_out(ctx, src, src_rows * src_cols, out, target_rows * target_cols,
                        [&](const float *ds, float *dout, cudaStream_t s) {
                            return launch_gather<MODE_RESAMPLE>(ctx, ds, src_rows, src_cols, dout,
                                                                target_rows, target_cols, gp, s);
                        });
}

}  // extern "C"
