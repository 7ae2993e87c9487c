// api.cu -- C-ABI entry points of libastroburst_cuda.so: context, memory helpers and the
// host-buffer / device-buffer stack calls.  See include/astroburst_cuda.h for the contract.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <climits>
#include <cmath>

#include "ab_internal.h"

ab_status ab_scratch_reserve(ab_ctx *ctx, int slot, size_t bytes) {
    if (ctx->d_scratch_bytes[slot] >= bytes) return AB_OK;
    if (ctx->d_scratch[slot]) {
        AB_CUDA(ctx, cudaFree(ctx->d_scratch[slot]));
        ctx->d_scratch[slot] = nullptr;
        ctx->d_scratch_bytes[slot] = 0;
    }
    cudaError_t e = cudaMalloc(&ctx->d_scratch[slot], bytes);
    if (e != cudaSuccess)
        return ab_fail(ctx, AB_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes,
                       cudaGetErrorString(e));
    ctx->d_scratch_bytes[slot] = bytes;
    return AB_OK;
}

extern "C" {

int ab_abi_version(void) { return AB_ABI_VERSION; }

ab_status ab_ctx_create(int device, ab_ctx **out) {
    if (!out) return AB_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return AB_ERR_CUDA;
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) return AB_ERR_CUDA;
    }
    if (device >= count) return AB_ERR_INVALID;
    ab_ctx *ctx = new (std::nothrow) ab_ctx();
    if (!ctx) return AB_ERR_NOMEM;
    ctx->device = device;
    DeviceGuard g(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return AB_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    {
        const char *e = getenv("AB_STACK_VARIANT");
        ctx->stack_variant = e ? atoi(e) : 0;
        e = getenv("AB_STACK_FORCE_TMA1D");
        ctx->force_tma1d = e && e[0] == '1';
        e = getenv("AB_STATS_HOST_WALKS");
        ctx->stats_host_walks = e && e[0] == '1';
        e = getenv("AB_STACK_MIX");
        ctx->stack_mix = e ? atoi(e) : 0;
    }
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < AB_NUM_COPY_STREAMS; ++i) {
        ok = cudaStreamCreateWithFlags(&ctx->copy_streams[i], cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->copy_events[i], cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaMalloc(&ctx->d_small, ab_ctx::SMALL_BYTES) == cudaSuccess;
    ok = ok && cudaMemset(ctx->d_small, 0, ab_ctx::SMALL_BYTES) == cudaSuccess;  // peer barrier flags start at 0
    ok = ok && cudaMalloc(&ctx->d_work_counters, ab_ctx::WORK_COUNTERS * sizeof(ab_ctx::StackSlot)) == cudaSuccess;
    ok = ok && cudaMemset(ctx->d_work_counters, 0, ab_ctx::WORK_COUNTERS * sizeof(ab_ctx::StackSlot)) == cudaSuccess;
    ok = ok && cudaMallocHost(&ctx->h_small, ab_ctx::SMALL_BYTES) == cudaSuccess;
    if (!ok) {
        ab_ctx_destroy(ctx);
        return AB_ERR_CUDA;
    }
    *out = ctx;
    return AB_OK;
}

void ab_ctx_destroy(ab_ctx *ctx) {
    if (!ctx) return;
    {
        DeviceGuard g(ctx->device);
        cudaDeviceSynchronize();
        for (int i = 0; i < AB_NUM_COPY_STREAMS; ++i) {
            if (ctx->d_scratch[i]) cudaFree(ctx->d_scratch[i]);
            if (ctx->copy_events[i]) cudaEventDestroy(ctx->copy_events[i]);
            if (ctx->copy_streams[i]) cudaStreamDestroy(ctx->copy_streams[i]);
        }
        ab_phasecorr_release(ctx);
        ab_peer_disconnect(ctx);
        if (ctx->d_small) cudaFree(ctx->d_small);
        if (ctx->d_work_counters) cudaFree(ctx->d_work_counters);
        if (ctx->h_small) cudaFreeHost(ctx->h_small);
        if (ctx->stream) cudaStreamDestroy(ctx->stream);
    }
    delete ctx;
}

const char *ab_last_error(ab_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }
int ab_ctx_device(ab_ctx *ctx) { return ctx ? ctx->device : -1; }
uint64_t ab_ctx_launch_count(ab_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

ab_status ab_ctx_synchronize(ab_ctx *ctx) {
    if (!ctx) return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < AB_NUM_COPY_STREAMS; ++i)
        AB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_streams[i]));
    return AB_OK;
}

// cudaDeviceEnablePeerAccess for this context's device (several GPUs driven from one process;
// across processes ab_ipc_open's lazy peer access does it).
ab_status ab_ctx_enable_peer_access(ab_ctx *ctx, int peer_device) {
    if (!ctx) return AB_ERR_INVALID;
    if (peer_device == ctx->device) return AB_OK;
    DeviceGuard g(ctx->device);
    int can = 0;
    AB_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
    if (!can)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "device %d cannot access device %d memory", ctx->device,
                       peer_device);
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        (void)cudaGetLastError();
        return AB_OK;
    }
    AB_CUDA(ctx, e);
    return AB_OK;
}

// CUDA IPC import under THIS context's device: cudaIpcMemLazyEnablePeerAccess makes the runtime
// set up peer access to the exporting GPU, so kernels of this context can load from the mapping
// (shard.cu: frame-sharded stacks read the owners' HBM over NVLink).
ab_status ab_ipc_open(ab_ctx *ctx, const void *handle64, void **dev_ptr) {
    if (!ctx || !handle64 || !dev_ptr) return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(&h, handle64, sizeof(h));
    *dev_ptr = nullptr;
    AB_CUDA(ctx, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return AB_OK;
}
ab_status ab_ipc_close(ab_ctx *ctx, void *dev_ptr) {
    if (!ctx) return AB_ERR_INVALID;
    if (!dev_ptr) return AB_OK;
    DeviceGuard g(ctx->device);
    AB_CUDA(ctx, cudaIpcCloseMemHandle(dev_ptr));
    return AB_OK;
}

ab_status ab_host_alloc(size_t bytes, void **out) {
    if (!out) return AB_ERR_INVALID;
    return cudaMallocHost(out, bytes ? bytes : 1) == cudaSuccess ? AB_OK : AB_ERR_NOMEM;
}
ab_status ab_host_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? AB_OK : AB_ERR_CUDA; }

ab_status ab_dev_alloc(ab_ctx *ctx, size_t bytes, void **out) {
    if (!ctx || !out) return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e != cudaSuccess)
        return ab_fail(ctx, AB_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes,
                       cudaGetErrorString(e));
    return AB_OK;
}
ab_status ab_dev_free(ab_ctx *ctx, void *p) {
    if (!ctx) return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    AB_CUDA(ctx, cudaFree(p));
    return AB_OK;
}
ab_status ab_memcpy_h2d(ab_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx) return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    AB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}
ab_status ab_memcpy_d2h(ab_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx) return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    AB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AB_OK;
}

// ------------------------------------------------------------------------------- stacking
// Input bytes per band of the host-buffer pipelines (AB_BAND_MB overrides, for measurements).
// Measured on B200 / PCIe Gen5 with 32 pinned 4096^2 frames: 32 MiB bands 342 Mpix/s, 96 MiB 395,
// 192 MiB 407, 384 MiB 417, 768 MiB 420 -- the per-frame copies of a band (band / n_frames bytes
// each) want to be >= 10 MB; three slots of 384 MiB are 1.2 GB of scratch.
static size_t ab_band_bytes() {
    const char *e = std::getenv("AB_BAND_MB");
    const long mb = e ? std::atol(e) : 0;
    return (size_t)(mb > 0 && mb <= 4096 ? mb : 384) << 20;
}
static ab_status check_stack_args(ab_ctx *ctx, const void *frames, size_t n_frames,
                                  const ab_stack_config *cfg, const void *out,
                                  bool align_ok = false) {
    if (!ctx) return AB_ERR_INVALID;
    if (n_frames == 0) return ab_fail(ctx, AB_ERR_INVALID, "No images to stack");
    if (!frames || !cfg || !out) return ab_fail(ctx, AB_ERR_INVALID, "null argument");
    if (cfg->align && !align_ok)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED,
                       "align=true needs whole frames: use ab_stack_images (device bands cannot "
                       "be phase-correlated)");
    if (n_frames > AB_MAX_FRAMES)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "n_frames %zu > AB_MAX_FRAMES %d", n_frames,
                       AB_MAX_FRAMES);
    return AB_OK;
}

ab_status ab_stack_sigma_clip_dev(ab_ctx *ctx, const float *const *d_frames, size_t n_frames,
                                  size_t npix, const ab_stack_config *cfg, float *d_out,
                                  uint64_t *d_rejected, void *stream) {
    ab_status st = check_stack_args(ctx, d_frames, n_frames, cfg, d_out);
    if (st != AB_OK) return st;
    if (!d_rejected) return ab_fail(ctx, AB_ERR_INVALID, "null d_rejected");
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    FrameTable ft;
    for (size_t f = 0; f < n_frames; ++f) {
        if (!d_frames[f]) return ab_fail(ctx, AB_ERR_INVALID, "null frame pointer %zu", f);
        ft.p[f] = d_frames[f];
    }
    for (size_t f = n_frames; f < AB_MAX_FRAMES; ++f) ft.p[f] = nullptr;
    cudaStream_t s = (cudaStream_t)stream;  // NULL = CUDA legacy default stream
    return ab_stack_launch(ctx, ft, (int)n_frames, npix, cfg->sigma_low, cfg->sigma_high,
                           cfg->max_iterations, d_out, (unsigned long long *)d_rejected, s);
}

// Host-buffer call: the image is cut into row bands; band k's frames are copied H2D on copy
// stream k % 3 while band k-1 is being combined and band k-2's result goes back D2H, so with
// pinned caller memory the call runs at PCIe speed with the kernel hidden behind the copies.
ab_status ab_stack_sigma_clip(ab_ctx *ctx, const float *const *frames, size_t n_frames,
                              size_t rows, size_t cols, const ab_stack_config *cfg, float *out,
                              uint64_t *rejected_pixels) {
    ab_status st = check_stack_args(ctx, frames, n_frames, cfg, out, true);
    if (st != AB_OK) return st;
    for (size_t f = 0; f < n_frames; ++f)
        if (!frames[f]) return ab_fail(ctx, AB_ERR_INVALID, "null frame pointer %zu", f);
    if (cfg->align)
        return ab_stack_images(ctx, frames, n_frames, rows, cols, cfg, out, rejected_pixels, nullptr);
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    if (rejected_pixels) *rejected_pixels = 0;
    const size_t npix = rows * cols;
    if (npix == 0) return AB_OK;

    // band size: ~384 MiB of input per band (ab_band_bytes), whole rows, at least one row
    const size_t target_px = std::max<size_t>(ab_band_bytes() / (n_frames * sizeof(float)), 1);
    size_t band_rows = std::max<size_t>(target_px / cols, 1);
    band_rows = std::min(band_rows, rows);
    const size_t band_px_max = band_rows * cols;
    const size_t n_bands = (rows + band_rows - 1) / band_rows;
    const int n_slots = (int)std::min<size_t>(AB_NUM_COPY_STREAMS, n_bands);

    // per slot: n_frames input planes + 1 output plane (256-byte aligned) ; counters in d_small
    const size_t plane_bytes = ((band_px_max * sizeof(float) + 255) / 256) * 256;
    for (int s = 0; s < n_slots; ++s) {
        st = ab_scratch_reserve(ctx, s, plane_bytes * (n_frames + 1));
        if (st != AB_OK) return st;
    }
    unsigned long long *d_rej = reinterpret_cast<unsigned long long *>(ctx->d_small);
    unsigned long long *h_rej = reinterpret_cast<unsigned long long *>(ctx->h_small);
    if (n_bands * sizeof(unsigned long long) > 4096)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "too many bands (%zu)", n_bands);

    for (size_t b = 0; b < n_bands; ++b) {
        const int slot = (int)(b % n_slots);
        cudaStream_t s = ctx->copy_streams[slot];
        const size_t r0 = b * band_rows;
        const size_t nr = std::min(band_rows, rows - r0);
        const size_t off = r0 * cols, px = nr * cols;
        char *base = static_cast<char *>(ctx->d_scratch[slot]);
        FrameTable ft;
        for (size_t f = 0; f < n_frames; ++f) {
            float *d = reinterpret_cast<float *>(base + f * plane_bytes);
            AB_CUDA(ctx, cudaMemcpyAsync(d, frames[f] + off, px * sizeof(float),
                                         cudaMemcpyHostToDevice, s));
            ft.p[f] = d;
        }
        for (size_t f = n_frames; f < AB_MAX_FRAMES; ++f) ft.p[f] = nullptr;
        float *d_out = reinterpret_cast<float *>(base + n_frames * plane_bytes);
        st = ab_stack_launch(ctx, ft, (int)n_frames, px, cfg->sigma_low, cfg->sigma_high,
                             cfg->max_iterations, d_out, d_rej + b, s);
        if (st != AB_OK) return st;
        AB_CUDA(ctx, cudaMemcpyAsync(out + off, d_out, px * sizeof(float),
                                     cudaMemcpyDeviceToHost, s));
        AB_CUDA(ctx, cudaMemcpyAsync(h_rej + b, d_rej + b, sizeof(unsigned long long),
                                     cudaMemcpyDeviceToHost, s));
    }
    for (int s = 0; s < n_slots; ++s) AB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_streams[s]));
    uint64_t total = 0;
    for (size_t b = 0; b < n_bands; ++b) total += h_rej[b];
    if (rejected_pixels) *rejected_pixels = total;
    return AB_OK;
}

// The same call on FITS data units as they sit in the files: every band crosses PCIe in the file's
// sample type (2 bytes per pixel for BITPIX 16 instead of the 4 of the decoded f32 plane) and
// decode_pixels (reader.rs:37-101) runs on the device right behind the copy, into the plane the
// stack kernel reads.  Same banding / three-stream pipeline as ab_stack_sigma_clip.
ab_status ab_stack_fits_units(ab_ctx *ctx, const uint8_t *const *units, const int32_t *bitpix,
                              const double *bscale, const double *bzero, size_t n_frames, size_t rows,
                              size_t cols, const ab_stack_config *cfg, float *out,
                              uint64_t *rejected_pixels) {
    ab_status st = check_stack_args(ctx, units, n_frames, cfg, out);
    if (st != AB_OK) return st;
    if (!bitpix || !bscale || !bzero) return ab_fail(ctx, AB_ERR_INVALID, "null argument");
    size_t max_bpp = 0;
    for (size_t f = 0; f < n_frames; ++f) {
        if (!units[f]) return ab_fail(ctx, AB_ERR_INVALID, "null data unit %zu", f);
        const size_t bpp = ab_fits_bytes_per_pixel(bitpix[f]);
        if (!bpp) return ab_fail(ctx, AB_ERR_UNSUPPORTED, "BITPIX %d (frame %zu)", (int)bitpix[f], f);
        max_bpp = std::max(max_bpp, bpp);
    }
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    if (rejected_pixels) *rejected_pixels = 0;
    const size_t npix = rows * cols;
    if (npix == 0) return AB_OK;

    const size_t target_px = std::max<size_t>(ab_band_bytes() / (n_frames * sizeof(float)), 1);
    size_t band_rows = std::min(std::max<size_t>(target_px / cols, 1), rows);
    const size_t band_px_max = band_rows * cols;
    const size_t n_bands = (rows + band_rows - 1) / band_rows;
    const int n_slots = (int)std::min<size_t>(AB_NUM_COPY_STREAMS, n_bands);
    // per slot: n_frames decoded planes + 1 output plane + two raw staging planes (frame f's copy
    // lands in staging f % 2 while frame f-1 is being decoded out of the other one)
    const size_t plane_bytes = ((band_px_max * sizeof(float) + 255) / 256) * 256;
    const size_t raw_bytes = ((band_px_max * max_bpp + 255) / 256) * 256;
    for (int s = 0; s < n_slots; ++s) {
        st = ab_scratch_reserve(ctx, s, plane_bytes * (n_frames + 1) + 2 * raw_bytes);
        if (st != AB_OK) return st;
    }
    unsigned long long *d_rej = reinterpret_cast<unsigned long long *>(ctx->d_small);
    unsigned long long *h_rej = reinterpret_cast<unsigned long long *>(ctx->h_small);
    if (n_bands * sizeof(unsigned long long) > 4096)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "too many bands (%zu)", n_bands);

    for (size_t b = 0; b < n_bands; ++b) {
        const int slot = (int)(b % n_slots);
        cudaStream_t s = ctx->copy_streams[slot];
        const size_t r0 = b * band_rows;
        const size_t nr = std::min(band_rows, rows - r0);
        const size_t off = r0 * cols, px = nr * cols;
        char *base = static_cast<char *>(ctx->d_scratch[slot]);
        unsigned char *raw = reinterpret_cast<unsigned char *>(base + plane_bytes * (n_frames + 1));
        FrameTable ft;
        for (size_t f = 0; f < n_frames; ++f) {
            float *d = reinterpret_cast<float *>(base + f * plane_bytes);
            const size_t bpp = ab_fits_bytes_per_pixel(bitpix[f]);
            unsigned char *stage = raw + (f & 1) * raw_bytes;
            AB_CUDA(ctx, cudaMemcpyAsync(stage, units[f] + off * bpp, px * bpp, cudaMemcpyHostToDevice, s));
            st = ab_fits_decode_on(ctx, stage, px, bitpix[f], bscale[f], bzero[f], d, s);
            if (st != AB_OK) return st;
            ft.p[f] = d;
        }
        for (size_t f = n_frames; f < AB_MAX_FRAMES; ++f) ft.p[f] = nullptr;
        float *d_out = reinterpret_cast<float *>(base + n_frames * plane_bytes);
        st = ab_stack_launch(ctx, ft, (int)n_frames, px, cfg->sigma_low, cfg->sigma_high,
                             cfg->max_iterations, d_out, d_rej + b, s);
        if (st != AB_OK) return st;
        AB_CUDA(ctx, cudaMemcpyAsync(out + off, d_out, px * sizeof(float), cudaMemcpyDeviceToHost, s));
        AB_CUDA(ctx, cudaMemcpyAsync(h_rej + b, d_rej + b, sizeof(unsigned long long),
                                     cudaMemcpyDeviceToHost, s));
    }
    for (int s = 0; s < n_slots; ++s) AB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_streams[s]));
    uint64_t total = 0;
    for (size_t b = 0; b < n_bands; ++b) total += h_rej[b];
    if (rejected_pixels) *rejected_pixels = total;
    return AB_OK;
}

// stack_images with config.align honoured (combine.rs:118-147): frame 0 is the reference, every
// other frame is phase-correlated against it (align.rs:84-97) and shifted with the bicubic
// sampler (align.rs:36-57); offsets[i] = (round(dy), round(dx)) like combine.rs:139-141.  All
// frames stay resident in HBM (n_frames + 1 planes), so this path is for stacks that fit the GPU.
ab_status ab_stack_images(ab_ctx *ctx, const float *const *frames, size_t n_frames, size_t rows,
                          size_t cols, const ab_stack_config *cfg, float *out,
                          uint64_t *rejected_pixels, int32_t *offsets) {
    ab_status st = check_stack_args(ctx, frames, n_frames, cfg, out, true);
    if (st != AB_OK) return st;
    for (size_t f = 0; f < n_frames; ++f)
        if (!frames[f]) return ab_fail(ctx, AB_ERR_INVALID, "null frame pointer %zu", f);
    if (offsets)
        for (size_t f = 0; f < 2 * n_frames; ++f) offsets[f] = 0;
    if (!cfg->align) {
        ab_stack_config c = *cfg;
        return ab_stack_sigma_clip(ctx, frames, n_frames, rows, cols, &c, out, rejected_pixels);
    }
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    if (rejected_pixels) *rejected_pixels = 0;
    const size_t npix = rows * cols;
    if (npix == 0) return AB_OK;
    const size_t plane_bytes = ((npix * sizeof(float) + 255) / 256) * 256;
    // slot 1: aligned planes + output; slot 0: the raw frame being aligned (the phase correlation
    // of the host-buffer entry point is not used here, so slot 0 is free)
    if ((st = ab_scratch_reserve(ctx, 1, plane_bytes * (n_frames + 1))) != AB_OK) return st;
    if ((st = ab_scratch_reserve(ctx, 0, plane_bytes)) != AB_OK) return st;
    char *base = static_cast<char *>(ctx->d_scratch[1]);
    float *d_raw = static_cast<float *>(ctx->d_scratch[0]);
    cudaStream_t s = ctx->stream;
    FrameTable ft;
    for (size_t f = 0; f < AB_MAX_FRAMES; ++f)
        ft.p[f] = f < n_frames ? reinterpret_cast<float *>(base + f * plane_bytes) : nullptr;
    float *d_ref = reinterpret_cast<float *>(base);
    AB_CUDA(ctx, cudaMemcpyAsync(d_ref, frames[0], npix * sizeof(float), cudaMemcpyHostToDevice, s));
    for (size_t f = 1; f < n_frames; ++f) {
        AB_CUDA(ctx, cudaMemcpyAsync(d_raw, frames[f], npix * sizeof(float), cudaMemcpyHostToDevice, s));
        ab_phase_result pc;
        // align == 1: the reference's operand order (bit parity with combine.rs:118-147, including its
        // sign convention, see the header); align == 2: operands swapped, so the measured offset is
        // the one shift_image_subpixel needs to bring the frame ONTO the reference
        if ((st = (cfg->align == 2 ? ab_phase_correlate_on(ctx, d_raw, d_ref, rows, cols, s, &pc)
                                   : ab_phase_correlate_on(ctx, d_ref, d_raw, rows, cols, s, &pc))) != AB_OK)
            return st;
        float *d_al = reinterpret_cast<float *>(base + f * plane_bytes);
        if ((st = ab_shift_bicubic_dev(ctx, d_raw, rows, cols, pc.dy, pc.dx, d_al, s)) != AB_OK)
            return st;
        if (offsets) {  // `.round() as i32`: half away from zero, saturating, NaN -> 0
            auto r32 = [](double v) -> int32_t {
                const double r = std::round(v);
                if (!(r == r)) return 0;
                return r >= 2147483647.0 ? INT32_MAX : (r <= -2147483648.0 ? INT32_MIN : (int32_t)r);
            };
            offsets[2 * f] = r32(pc.dy);
            offsets[2 * f + 1] = r32(pc.dx);
        }
    }
    float *d_out = reinterpret_cast<float *>(base + n_frames * plane_bytes);
    unsigned long long *d_rej = reinterpret_cast<unsigned long long *>(ctx->d_small);
    unsigned long long *h_rej = reinterpret_cast<unsigned long long *>(ctx->h_small);
    st = ab_stack_launch(ctx, ft, (int)n_frames, npix, cfg->sigma_low, cfg->sigma_high,
                         cfg->max_iterations, d_out, d_rej, s);
    if (st != AB_OK) return st;
    AB_CUDA(ctx, cudaMemcpyAsync(out, d_out, npix * sizeof(float), cudaMemcpyDeviceToHost, s));
    AB_CUDA(ctx, cudaMemcpyAsync(h_rej, d_rej, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    AB_CUDA(ctx, cudaStreamSynchronize(s));
    if (rejected_pixels) *rejected_pixels = *h_rej;
    return AB_OK;
}

ab_status ab_sigma_clip_combine(ab_ctx *ctx, const float *values, size_t n, float sigma_low,
                                float sigma_high, uint32_t max_iter, float *mean_out,
                                uint32_t *rejected_out) {
    if (!ctx || !mean_out || !rejected_out) return AB_ERR_INVALID;
    if (n > AB_MAX_FRAMES)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "n %zu > AB_MAX_FRAMES %d", n, AB_MAX_FRAMES);
    *mean_out = 0.0f;
    *rejected_out = 0;
    if (n == 0) return AB_OK;  // combine.rs:21-23
    if (!values) return ab_fail(ctx, AB_ERR_INVALID, "null values");
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    // layout in d_small: [0,8) counter | [64, 64+4) out | [8192, ...) values
    char *d = static_cast<char *>(ctx->d_small);
    float *d_vals = reinterpret_cast<float *>(d + 8192);
    float *d_out = reinterpret_cast<float *>(d + 64);
    unsigned long long *d_rej = reinterpret_cast<unsigned long long *>(d);
    AB_CUDA(ctx, cudaMemcpyAsync(d_vals, values, n * sizeof(float), cudaMemcpyHostToDevice,
                                 ctx->stream));
    FrameTable ft;
    for (size_t f = 0; f < AB_MAX_FRAMES; ++f) ft.p[f] = f < n ? d_vals + f : nullptr;
    ab_status st = ab_stack_launch(ctx, ft, (int)n, 1, sigma_low, sigma_high, max_iter, d_out,
                                   d_rej, ctx->stream);
    if (st != AB_OK) return st;
    unsigned long long rej = 0;
    AB_CUDA(ctx, cudaMemcpyAsync(mean_out, d_out, sizeof(float), cudaMemcpyDeviceToHost,
                                 ctx->stream));
    AB_CUDA(ctx, cudaMemcpyAsync(&rej, d_rej, sizeof(rej), cudaMemcpyDeviceToHost, ctx->stream));
    AB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *rejected_out = (uint32_t)rej;
    return AB_OK;
}

}  // extern "C"
