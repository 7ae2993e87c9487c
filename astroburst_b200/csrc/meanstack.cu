// meanstack.cu -- sigma_clipped_mean_stack of the batch calibration pipeline.
//
// Reference: src-tauri/src/core/imaging/calibration_pipeline.rs:317-378.  NOT the stack of
// combine.rs (stack_kernels.cuh): here every sample takes part (no finite filter; the total order
// of math/median.rs:4-13 puts NaN last), the centre / scale of EVERY iteration are the upper
// median and the upper-median MAD, sigma = (mad as f64 * 1.4826) as f32 with `sigma < 1e-10`
// ending the loop, the clip is the strict f32 test -sigma_low < (v - median) / sigma < sigma_high,
// rejections are counted PER FRAME, and the result is the f32 sum of the survivors in frame order
// divided by their count (f32) -- 0 when nothing survives.
//
// One thread = one pixel, P = frame count padded to 8/16/32/64/128/256.  Samples stay in frame
// order in registers next to an alive mask; each iteration sorts order-preserving integer keys
// (dead samples = all ones, canonical NaN just below) with the compare-exchange network of the
// stack kernels, once for the median and once for the deviations.
#include <algorithm>
#include <vector>

#include "stack_kernels.cuh"

namespace {

constexpr int MS_THREADS = 128;
constexpr unsigned KEY_DEAD = 0xffffffffu;
constexpr unsigned KEY_NAN = 0xffc00000u;  // every NaN: after +inf (0xff800000), before the dead

// f32_cmp order (NaN last, -0 == +0 up to their order, which never changes a value used below)
__device__ __forceinline__ unsigned key_of(float v) {
    if (v != v) return KEY_NAN;
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float value_of(unsigned k) {
    if (k == KEY_NAN) return CUDART_NAN_F;
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ void ce_u32(unsigned &a, unsigned &b) {
    const unsigned lo = min(a, b);
    b = max(a, b);
    a = lo;
}
template <int P, int... I>
__device__ __forceinline__ void sort_keys_impl(unsigned (&k)[P], std::integer_sequence<int, I...>) {
    (ce_u32(k[abk::OemHolder<P>::net.pr[I].a], k[abk::OemHolder<P>::net.pr[I].b]), ...);
}
// P <= 64: the register network; 128 / 256: a rolled bitonic sort on the (local-memory) array
template <int P>
__device__ __forceinline__ void sort_keys(unsigned (&k)[P]) {
    if constexpr (P <= 64) {
        sort_keys_impl<P>(k, std::make_integer_sequence<int, abk::OemHolder<P>::net.count>{});
    } else {
#pragma unroll 1
        for (int size = 2; size <= P; size <<= 1)
#pragma unroll 1
            for (int stride = size >> 1; stride > 0; stride >>= 1)
#pragma unroll 1
                for (int i = 0; i < P; ++i) {
                    const int j = i ^ stride;
                    if (j > i) {
                        const bool up = (i & size) == 0;
                        const unsigned a = k[i], b = k[j];
                        if ((a > b) == up) {
                            k[i] = b;
                            k[j] = a;
                        }
                    }
                }
    }
}
// k[idx] for a runtime idx (P <= 64: without forcing the array out of registers)
template <int P>
__device__ __forceinline__ unsigned pick(const unsigned (&k)[P], int idx) {
    if constexpr (P <= 64) {
        unsigned r = 0;
#pragma unroll
        for (int j = 0; j < P; ++j) r = j == idx ? k[j] : r;
        return r;
    } else {
        return k[idx];
    }
}

template <int P>
__global__ void __launch_bounds__(MS_THREADS)
mean_stack_kernel(const __grid_constant__ FrameTable ft, int n, size_t npix, float sigma_low,
                  float sigma_high, unsigned max_iter, float *__restrict__ out,
                  unsigned long long *__restrict__ rejection_counts) {
    __shared__ unsigned rej_sh[P];
    for (int j = threadIdx.x; j < P; j += MS_THREADS) rej_sh[j] = 0;
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * MS_THREADS;
    for (size_t p = (size_t)blockIdx.x * MS_THREADS + threadIdx.x; p < npix; p += stride) {
        float v[P];
        bool alive[P];
#pragma unroll
        for (int j = 0; j < P; ++j) {
            alive[j] = j < n;
            v[j] = j < n ? __ldcs(ft.p[j] + p) : 0.0f;
        }
        int len = n;
        for (unsigned it = 0; it < max_iter; ++it) {
            if (len < 3) break;
            const int mid = len >> 1;
            unsigned k[P];
#pragma unroll
            for (int j = 0; j < P; ++j) k[j] = alive[j] ? key_of(v[j]) : KEY_DEAD;
            sort_keys<P>(k);
            const float median = value_of(pick<P>(k, mid));
#pragma unroll
            for (int j = 0; j < P; ++j) k[j] = alive[j] ? key_of(fabsf(v[j] - median)) : KEY_DEAD;
            sort_keys<P>(k);
            const float mad = value_of(pick<P>(k, mid));
            const float sigma = (float)((double)mad * abk::MAD_TO_SIGMA);
            if (sigma < 1e-10f) break;
            int removed = 0;
#pragma unroll
            for (int j = 0; j < P; ++j) {
                if (alive[j]) {
                    const float z = (v[j] - median) / sigma;
                    const bool keep = z > -sigma_low && z < sigma_high;
                    if (!keep) {
                        alive[j] = false;
                        ++removed;
                        atomicAdd(&rej_sh[j], 1u);
                    }
                }
            }
            if (removed == 0) break;
            len -= removed;
        }
        float sum = 0.0f;  // frame order, f32, like vals.iter().sum::<f32>()
#pragma unroll
        for (int j = 0; j < P; ++j)
            if (alive[j]) sum += v[j];
        __stcs(out + p, len == 0 ? 0.0f : sum / (float)len);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += MS_THREADS)
        if (rej_sh[j]) atomicAdd(rejection_counts + j, (unsigned long long)rej_sh[j]);
}

template <int P>
ab_status launch_mean_stack(ab_ctx *ctx, const FrameTable &ft, int n, size_t npix, float sl, float sh,
                            unsigned max_iter, float *d_out, unsigned long long *d_counts,
                            cudaStream_t s) {
    size_t blocks = (npix + MS_THREADS - 1) / MS_THREADS;
    const size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    mean_stack_kernel<P><<<(unsigned)blocks, MS_THREADS, 0, s>>>(ft, n, npix, sl, sh, max_iter, d_out,
                                                                 d_counts);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

ab_status mean_stack_on(ab_ctx *ctx, const FrameTable &ft, int n, size_t npix, float sl, float sh,
                        unsigned max_iter, float *d_out, unsigned long long *d_counts, cudaStream_t s) {
    AB_CUDA(ctx, cudaMemsetAsync(d_counts, 0, (size_t)n * sizeof(unsigned long long), s));
    if (npix == 0) return AB_OK;
    if (n <= 8) return launch_mean_stack<8>(ctx, ft, n, npix, sl, sh, max_iter, d_out, d_counts, s);
    if (n <= 16) return launch_mean_stack<16>(ctx, ft, n, npix, sl, sh, max_iter, d_out, d_counts, s);
    if (n <= 32) return launch_mean_stack<32>(ctx, ft, n, npix, sl, sh, max_iter, d_out, d_counts, s);
    if (n <= 64) return launch_mean_stack<64>(ctx, ft, n, npix, sl, sh, max_iter, d_out, d_counts, s);
    if (n <= 128) return launch_mean_stack<128>(ctx, ft, n, npix, sl, sh, max_iter, d_out, d_counts, s);
    return launch_mean_stack<256>(ctx, ft, n, npix, sl, sh, max_iter, d_out, d_counts, s);
}

}  // namespace

extern "C" {

ab_status ab_sigma_clipped_mean_stack_dev(ab_ctx *ctx, const float *const *d_frames, size_t n_frames,
                                          size_t npix, float sigma_low, float sigma_high,
                                          uint32_t max_iterations, float *d_out,
                                          uint64_t *d_rejection_counts, void *stream) {
    if (!ctx) return AB_ERR_INVALID;
    if (n_frames == 0) return ab_fail(ctx, AB_ERR_INVALID, "no frames");
    if (!d_frames || !d_rejection_counts || (npix && !d_out)) return ab_fail(ctx, AB_ERR_INVALID, "null pointer");
    if (n_frames > AB_MAX_FRAMES)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "n_frames %zu > AB_MAX_FRAMES %d", n_frames, AB_MAX_FRAMES);
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    FrameTable ft;
    for (size_t f = 0; f < AB_MAX_FRAMES; ++f) ft.p[f] = f < n_frames ? d_frames[f] : nullptr;
    for (size_t f = 0; f < n_frames; ++f)
        if (!ft.p[f]) return ab_fail(ctx, AB_ERR_INVALID, "null frame pointer %zu", f);
    return mean_stack_on(ctx, ft, (int)n_frames, npix, sigma_low, sigma_high, max_iterations, d_out,
                         reinterpret_cast<unsigned long long *>(d_rejection_counts), (cudaStream_t)stream);
}

// Host-buffer call: row bands of ~96 MiB of input, copies of band b+1 overlap the kernel of band b.
ab_status ab_sigma_clipped_mean_stack(ab_ctx *ctx, const float *const *frames, size_t n_frames,
                                      size_t rows, size_t cols, float sigma_low, float sigma_high,
                                      uint32_t max_iterations, float *out, uint64_t *rejection_counts) {
    if (!ctx) return AB_ERR_INVALID;
    if (n_frames == 0) return ab_fail(ctx, AB_ERR_INVALID, "no frames");
    if (!frames || !rejection_counts || !out) return ab_fail(ctx, AB_ERR_INVALID, "null pointer");
    if (n_frames > AB_MAX_FRAMES)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "n_frames %zu > AB_MAX_FRAMES %d", n_frames, AB_MAX_FRAMES);
    for (size_t f = 0; f < n_frames; ++f) {
        if (!frames[f]) return ab_fail(ctx, AB_ERR_INVALID, "null frame pointer %zu", f);
        rejection_counts[f] = 0;
    }
    const size_t npix = rows * cols;
    if (npix == 0) return AB_OK;
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    const size_t target_px = std::max<size_t>((96u << 20) / ((n_frames + 1) * sizeof(float)), 1);
    const size_t band_rows = std::min(std::max<size_t>(target_px / cols, 1), rows);
    const size_t n_bands = (rows + band_rows - 1) / band_rows;
    const int n_slots = (int)std::min<size_t>(AB_NUM_COPY_STREAMS, n_bands);
    const size_t plane_bytes = ((band_rows * cols * sizeof(float) + 255) / 256) * 256;
    const size_t cnt_bytes = AB_MAX_FRAMES * sizeof(unsigned long long);
    for (int s = 0; s < n_slots; ++s) {
        ab_status st = ab_scratch_reserve(ctx, s, plane_bytes * (n_frames + 1) + cnt_bytes);
        if (st != AB_OK) return st;
    }
    std::vector<unsigned long long> h_cnt(n_frames);
    for (size_t b = 0; b < n_bands; ++b) {
        const int slot = (int)(b % n_slots);
        cudaStream_t s = ctx->copy_streams[slot];
        const size_t r0 = b * band_rows, nr = std::min(band_rows, rows - r0);
        const size_t off = r0 * cols, px = nr * cols;
        char *base = static_cast<char *>(ctx->d_scratch[slot]);
        FrameTable ft;
        for (size_t f = 0; f < AB_MAX_FRAMES; ++f) ft.p[f] = nullptr;
        for (size_t f = 0; f < n_frames; ++f) {
            float *d = reinterpret_cast<float *>(base + f * plane_bytes);
            AB_CUDA(ctx, cudaMemcpyAsync(d, frames[f] + off, px * sizeof(float), cudaMemcpyHostToDevice, s));
            ft.p[f] = d;
        }
        float *d_out = reinterpret_cast<float *>(base + n_frames * plane_bytes);
        unsigned long long *d_cnt = reinterpret_cast<unsigned long long *>(base + (n_frames + 1) * plane_bytes);
        ab_status st = mean_stack_on(ctx, ft, (int)n_frames, px, sigma_low, sigma_high, max_iterations,
                                     d_out, d_cnt, s);
        if (st != AB_OK) return st;
        AB_CUDA(ctx, cudaMemcpyAsync(out + off, d_out, px * sizeof(float), cudaMemcpyDeviceToHost, s));
        AB_CUDA(ctx, cudaMemcpyAsync(h_cnt.data(), d_cnt, n_frames * sizeof(unsigned long long),
                                     cudaMemcpyDeviceToHost, s));
        AB_CUDA(ctx, cudaStreamSynchronize(s));  // counters are per band: add them up on the host
        for (size_t f = 0; f < n_frames; ++f) rejection_counts[f] += h_cnt[f];
    }
    return AB_OK;
}

}  // extern "C"
