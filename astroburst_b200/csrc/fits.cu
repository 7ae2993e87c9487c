// fits.cu -- FITS data-unit decode / encode on the device: the wire format on
// either side of the path, so frames can go disk -> pinned bytes -> GPU without a host decode
// (and 16-bit camera frames cross PCIe at half the bytes of their f32 form).
//
// Reference (src-tauri/src/infra/fits/):
//   reader.rs:37-101   is_identity_scaling, decode_pixels: big-endian BITPIX 8 / 16 / 32 / -32 / -64
//                      -> f32, `(v as f64 * bscale + bzero) as f32` unless the scaling is identity
//   writer.rs:82-136   write_f32_slice_as_be, write_i16_slice_as_be
//                      (`((val as f64 - bzero) / bscale).clamp(i16::MIN, i16::MAX).round() as i16`),
//                      write_f64_slice_as_be
// Header parsing, file I/O and block padding stay with the caller (infra, out of scope).
#include <algorithm>
#include <cmath>

#include "ab_internal.h"

namespace {

constexpr int FITS_THREADS = 256;

__device__ __forceinline__ unsigned bswap32(unsigned v) { return __byte_perm(v, 0, 0x0123); }

template <int BITPIX>
__global__ void __launch_bounds__(FITS_THREADS)
decode_kernel(const unsigned char *__restrict__ data, size_t n, double bscale, double bzero,
              bool identity, float *__restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float r;
        if constexpr (BITPIX == 8) {
            const unsigned char b = data[i];
            r = identity ? (float)b : (float)((double)b * bscale + bzero);
        } else if constexpr (BITPIX == 16) {
            const unsigned short w = reinterpret_cast<const unsigned short *>(data)[i];
            const short v = (short)(unsigned short)((w >> 8) | (w << 8));
            r = identity ? (float)v : (float)((double)v * bscale + bzero);
        } else if constexpr (BITPIX == 32) {
            const int v = (int)bswap32(reinterpret_cast<const unsigned *>(data)[i]);
            r = identity ? (float)v : (float)((double)v * bscale + bzero);
        } else if constexpr (BITPIX == -32) {
            const float v = __uint_as_float(bswap32(reinterpret_cast<const unsigned *>(data)[i]));
            r = identity ? v : (float)((double)v * bscale + bzero);
        } else {  // -64
            const uint2 w = reinterpret_cast<const uint2 *>(data)[i];
            const double v = __hiloint2double((int)bswap32(w.x), (int)bswap32(w.y));
            r = identity ? (float)v : (float)(v * bscale + bzero);
        }
        out[i] = r;
    }
}

template <int BITPIX>
__global__ void __launch_bounds__(FITS_THREADS)
encode_kernel(const float *__restrict__ data, size_t n, double bzero, double bscale,
              unsigned char *__restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float val = data[i];
        if constexpr (BITPIX == -32) {
            reinterpret_cast<unsigned *>(out)[i] = bswap32(__float_as_uint(val));
        } else if constexpr (BITPIX == 16) {
            const double physical = ((double)val - bzero) / bscale;
            // f64::clamp keeps NaN; `.round() as i16` saturates and maps NaN to 0
            double c = physical < -32768.0 ? -32768.0 : (physical > 32767.0 ? 32767.0 : physical);
            c = round(c);
            const short v = (c == c) ? (short)c : (short)0;
            const unsigned short w = (unsigned short)v;
            reinterpret_cast<unsigned short *>(out)[i] = (unsigned short)((w >> 8) | (w << 8));
        } else {  // -64
            const double d = (double)val;
            const unsigned hi = (unsigned)__double2hiint(d), lo = (unsigned)__double2loint(d);
            reinterpret_cast<uint2 *>(out)[i] = make_uint2(bswap32(hi), bswap32(lo));
        }
    }
}

unsigned fits_blocks(ab_ctx *ctx, size_t n) {
    const size_t b = (n + FITS_THREADS - 1) / FITS_THREADS, cap = (size_t)ctx->sm_count * 16;
    return (unsigned)std::max<size_t>(1, std::min(b, cap));
}

int bytes_per_pixel(int bitpix) {
    switch (bitpix) {
        case 8: return 1;
        case 16: return 2;
        case 32: case -32: return 4;
        case -64: return 8;
        default: return 0;
    }
}

ab_status decode_on(ab_ctx *ctx, const unsigned char *d, size_t n, int bitpix, double bscale,
                    double bzero, float *out, cudaStream_t s) {
    const bool identity = std::fabs(bscale - 1.0) < 1e-15 && std::fabs(bzero) < 1e-15;  // reader.rs:37-40
    const unsigned g = fits_blocks(ctx, n);
    switch (bitpix) {
        case 8: decode_kernel<8><<<g, FITS_THREADS, 0, s>>>(d, n, bscale, bzero, identity, out); break;
        case 16: decode_kernel<16><<<g, FITS_THREADS, 0, s>>>(d, n, bscale, bzero, identity, out); break;
        case 32: decode_kernel<32><<<g, FITS_THREADS, 0, s>>>(d, n, bscale, bzero, identity, out); break;
        case -32: decode_kernel<-32><<<g, FITS_THREADS, 0, s>>>(d, n, bscale, bzero, identity, out); break;
        case -64: decode_kernel<-64><<<g, FITS_THREADS, 0, s>>>(d, n, bscale, bzero, identity, out); break;
        default: return ab_fail(ctx, AB_ERR_UNSUPPORTED, "BITPIX %d", bitpix);
    }
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

ab_status encode_on(ab_ctx *ctx, const float *d, size_t n, int bitpix, double bzero, double bscale,
                    unsigned char *out, cudaStream_t s) {
    const unsigned g = fits_blocks(ctx, n);
    switch (bitpix) {
        case 16: encode_kernel<16><<<g, FITS_THREADS, 0, s>>>(d, n, bzero, bscale, out); break;
        case -64: encode_kernel<-64><<<g, FITS_THREADS, 0, s>>>(d, n, bzero, bscale, out); break;
        case -32: encode_kernel<-32><<<g, FITS_THREADS, 0, s>>>(d, n, bzero, bscale, out); break;
        default: return ab_fail(ctx, AB_ERR_UNSUPPORTED, "BITPIX %d (writer: 16, -32, -64)", bitpix);
    }
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

}  // namespace

ab_status ab_fits_decode_on(ab_ctx *ctx, const unsigned char *d, size_t n, int bitpix, double bscale,
                            double bzero, float *out, cudaStream_t s) {
    return decode_on(ctx, d, n, bitpix, bscale, bzero, out, s);
}
size_t ab_fits_bytes_per_pixel(int bitpix) { return (size_t)bytes_per_pixel(bitpix); }

extern "C" {

ab_status ab_fits_decode_dev(ab_ctx *ctx, const uint8_t *d_data, size_t n_pixels, int bitpix,
                             double bscale, double bzero, float *d_out, void *stream) {
    if (!ctx || (n_pixels && (!d_data || !d_out))) return AB_ERR_INVALID;
    if (!bytes_per_pixel(bitpix)) return ab_fail(ctx, AB_ERR_UNSUPPORTED, "BITPIX %d", bitpix);
    if (n_pixels == 0) return AB_OK;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    return decode_on(ctx, d_data, n_pixels, bitpix, bscale, bzero, d_out, (cudaStream_t)stream);
}

ab_status ab_fits_decode(ab_ctx *ctx, const uint8_t *data, size_t n_pixels, int bitpix, double bscale,
                         double bzero, float *out) {
    if (!ctx || (n_pixels && (!data || !out))) return AB_ERR_INVALID;
    const int bpp = bytes_per_pixel(bitpix);
    if (!bpp) return ab_fail(ctx, AB_ERR_UNSUPPORTED, "BITPIX %d", bitpix);
    if (n_pixels == 0) return AB_OK;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    const size_t in_b = (n_pixels * bpp + 255) / 256 * 256;
    ab_status st = ab_scratch_reserve(ctx, 0, in_b + n_pixels * sizeof(float));
    if (st != AB_OK) return st;
    unsigned char *d_in = static_cast<unsigned char *>(ctx->d_scratch[0]);
    float *d_out = reinterpret_cast<float *>(d_in + in_b);
    AB_CUDA(ctx, cudaMemcpyAsync(d_in, data, n_pixels * bpp, cudaMemcpyHostToDevice, ctx->stream));
    st = decode_on(ctx, d_in, n_pixels, bitpix, bscale, bzero, d_out, ctx->stream);
    if (st != AB_OK) return st;
    AB_CUDA(ctx, cudaMemcpyAsync(out, d_out, n_pixels * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}

ab_status ab_fits_encode_dev(ab_ctx *ctx, const float *d_data, size_t n_pixels, int bitpix,
                             double bzero, double bscale, uint8_t *d_out, void *stream) {
    if (!ctx || (n_pixels && (!d_data || !d_out))) return AB_ERR_INVALID;
    if (n_pixels == 0) return AB_OK;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    return encode_on(ctx, d_data, n_pixels, bitpix, bzero, bscale, d_out, (cudaStream_t)stream);
}

ab_status ab_fits_encode(ab_ctx *ctx, const float *data, size_t n_pixels, int bitpix, double bzero,
                         double bscale, uint8_t *out) {
    if (!ctx || (n_pixels && (!data || !out))) return AB_ERR_INVALID;
    if (bitpix != 16 && bitpix != -32 && bitpix != -64)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "BITPIX %d (writer: 16, -32, -64)", bitpix);
    if (n_pixels == 0) return AB_OK;
    const int bpp = bytes_per_pixel(bitpix);
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    const size_t in_b = (n_pixels * sizeof(float) + 255) / 256 * 256;
    ab_status st = ab_scratch_reserve(ctx, 0, in_b + n_pixels * bpp);
    if (st != AB_OK) return st;
    float *d_in = static_cast<float *>(ctx->d_scratch[0]);
    unsigned char *d_out = reinterpret_cast<unsigned char *>(d_in) + in_b;
    AB_CUDA(ctx, cudaMemcpyAsync(d_in, data, n_pixels * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    st = encode_on(ctx, d_in, n_pixels, bitpix, bzero, bscale, d_out, ctx->stream);
    if (st != AB_OK) return st;
    AB_CUDA(ctx, cudaMemcpyAsync(out, d_out, n_pixels * bpp, cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}

}  // extern "C"
