// orc_fits.cpp -- CPU oracle (TEST INFRASTRUCTURE, see oracle.h) for the FITS data-unit codec.
// Restates src-tauri/src/infra/fits/reader.rs:37-101 (is_identity_scaling, decode_pixels) and
// writer.rs:82-136 (write_f32/i16/f64_slice_as_be).  Rust `as` casts: f64 -> f32 rounds to nearest,
// f64 -> i16 saturates and maps NaN to 0; f64::round is half away from zero; f64::clamp keeps NaN.
#include <cmath>
#include <cstdint>
#include <cstring>

#include "oracle.h"

extern "C" {

int orc_fits_decode(const uint8_t *data, size_t n, int bitpix, double bscale, double bzero, float *out) {
    const bool identity = std::fabs(bscale - 1.0) < 1e-15 && std::fabs(bzero) < 1e-15;
    for (size_t i = 0; i < n; ++i) {
        switch (bitpix) {
            case 8: {
                const uint8_t b = data[i];
                out[i] = identity ? (float)b : (float)((double)b * bscale + bzero);
                break;
            }
            case 16: {
                const int16_t v = (int16_t)((uint16_t)data[2 * i] << 8 | data[2 * i + 1]);
                out[i] = identity ? (float)v : (float)((double)v * bscale + bzero);
                break;
            }
            case 32: {
                const uint32_t u = (uint32_t)data[4 * i] << 24 | (uint32_t)data[4 * i + 1] << 16 |
                                   (uint32_t)data[4 * i + 2] << 8 | data[4 * i + 3];
                const int32_t v = (int32_t)u;
                out[i] = identity ? (float)v : (float)((double)v * bscale + bzero);
                break;
            }
            case -32: {
                const uint32_t u = (uint32_t)data[4 * i] << 24 | (uint32_t)data[4 * i + 1] << 16 |
                                   (uint32_t)data[4 * i + 2] << 8 | data[4 * i + 3];
                float v;
                std::memcpy(&v, &u, 4);
                out[i] = identity ? v : (float)((double)v * bscale + bzero);
                break;
            }
            case -64: {
                uint64_t u = 0;
                for (int k = 0; k < 8; ++k) u = u << 8 | data[8 * i + k];
                double v;
                std::memcpy(&v, &u, 8);
                out[i] = identity ? (float)v : (float)(v * bscale + bzero);
                break;
            }
            default:
                return 1;  // the reference returns an empty Vec
        }
    }
    return 0;
}

int orc_fits_encode(const float *data, size_t n, int bitpix, double bzero, double bscale, uint8_t *out) {
    for (size_t i = 0; i < n; ++i) {
        const float val = data[i];
        if (bitpix == -32) {
            uint32_t u;
            std::memcpy(&u, &val, 4);
            for (int k = 0; k < 4; ++k) out[4 * i + k] = (uint8_t)(u >> (24 - 8 * k));
        } else if (bitpix == 16) {
            const double physical = ((double)val - bzero) / bscale;
            double c = physical < -32768.0 ? -32768.0 : (physical > 32767.0 ? 32767.0 : physical);
            c = std::round(c);
            const int16_t v = (c == c) ? (int16_t)c : (int16_t)0;
            out[2 * i] = (uint8_t)((uint16_t)v >> 8);
            out[2 * i + 1] = (uint8_t)((uint16_t)v & 0xff);
        } else if (bitpix == -64) {
            const double d = (double)val;
            uint64_t u;
            std::memcpy(&u, &d, 8);
            for (int k = 0; k < 8; ++k) out[8 * i + k] = (uint8_t)(u >> (56 - 8 * k));
        } else {
            return 1;
        }
    }
    return 0;
}

}  // extern "C"
