// stack.cu -- dispatcher of the sigma-clipped frame stack and the deep-stack kernel.
//
// Reference semantics: stack_images / sigma_clip_combine,
// src-tauri/src/core/stacking/combine.rs:14-92,160-182 (align=false core loop).
// The register-resident kernels (exact frame count P, see stack_kernels.cuh) are instantiated
// in stack_inst_*.cu for P = 1..32, 36, 40, 48, 56, 64; every other count <= 256 runs the
// deep-stack kernel below.
#include <cfloat>
#include <cmath>

#include <cuda.h>

#include "stack_kernels.cuh"

namespace abk {

// [P, npix] f32 view of the frame cube, box = box_px pixels x P frames, no swizzle: the tile lands
// in shared memory as stage[f][pixel], exactly what the per-thread column reads expect.
bool ab_make_stack_tmap(ab_ctx *ctx, const float *base, size_t npix, size_t plane_stride_elems,
                        int P, int box_px, TmapParam *out) {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    if (!ctx->encode_tmap_fn) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) !=
                cudaSuccess ||
            qres != cudaDriverEntryPointSuccess || !fn)
            return false;
        ctx->encode_tmap_fn = fn;
    }
    static_assert(sizeof(CUtensorMap) == sizeof(TmapParam), "CUtensorMap is 128 bytes");
    const cuuint64_t gdim[2] = {(cuuint64_t)npix, (cuuint64_t)P};
    const cuuint64_t gstride[1] = {(cuuint64_t)plane_stride_elems * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_px, (cuuint32_t)P};
    const cuuint32_t estride[2] = {1, 1};
    CUresult r = reinterpret_cast<encode_fn>(ctx->encode_tmap_fn)(
        reinterpret_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
        const_cast<float *>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

#define AB_DECL(P)                                                                             \
    extern template ab_status launch_tile<P>(ab_ctx *, const FrameTable &, size_t, float, float, \
                                             unsigned, float *, unsigned long long *, cudaStream_t);
AB_DECL(1) AB_DECL(2) AB_DECL(3) AB_DECL(4) AB_DECL(5) AB_DECL(6) AB_DECL(7) AB_DECL(8)
AB_DECL(9) AB_DECL(10) AB_DECL(11) AB_DECL(12) AB_DECL(13) AB_DECL(14) AB_DECL(15) AB_DECL(16)
AB_DECL(17) AB_DECL(18) AB_DECL(19) AB_DECL(20) AB_DECL(21) AB_DECL(22) AB_DECL(23) AB_DECL(24)
AB_DECL(25) AB_DECL(26) AB_DECL(27) AB_DECL(28) AB_DECL(29) AB_DECL(30) AB_DECL(31) AB_DECL(32)
AB_DECL(36) AB_DECL(40) AB_DECL(48) AB_DECL(56) AB_DECL(64)
#undef AB_DECL
}  // namespace abk

namespace {

// ---------------------------------------------------------------------------------------------
// Deep-stack path (frame counts without a register kernel, up to 256): one thread per pixel,
// samples in a shared-memory column smem[i * DEEP_THREADS + tid] (conflict-free), bitonic sort
// with runtime loops, then the dynamic-index algorithm.  Correct for any n; a warp-cooperative
// version is future work.
constexpr int DEEP_THREADS = 64;

__global__ void __launch_bounds__(DEEP_THREADS)
stack_deep_kernel(const __grid_constant__ FrameTable ft, int n_frames, int n_pad, size_t npix,
                  float sigma_low, float sigma_high, unsigned max_iter, float *__restrict__ out,
                  unsigned long long *__restrict__ rejected_total) {
    extern __shared__ float sm[];
    float *col = sm + threadIdx.x;
    const int stride = DEEP_THREADS;
    unsigned long long local_rej = 0;
    for (size_t base = (size_t)blockIdx.x * blockDim.x; base < npix;
         base += (size_t)gridDim.x * blockDim.x) {
        const size_t p = base + threadIdx.x;
        const bool active = p < npix;
        int n = 0;
        for (int f = 0; f < n_pad; ++f) {
            float v = CUDART_INF_F;
            if (active && f < n_frames) {
                v = __ldcs(ft.p[f] + p);
                const bool fin = fabsf(v) < CUDART_INF_F;
                n += fin ? 1 : 0;
                v = fin ? v : CUDART_INF_F;
            }
            col[f * stride] = v;
        }
        // bitonic sort, ascending, private column (no cross-thread hazards)
        for (int k = 2; k <= n_pad; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = 0; i < n_pad; ++i) {
                    const int l = i ^ j;
                    if (l > i) {
                        const float x = col[i * stride], y = col[l * stride];
                        const bool up = (i & k) == 0;
                        if (up ? (x > y) : (x < y)) {
                            col[i * stride] = y;
                            col[l * stride] = x;
                        }
                    }
                }
            }
        }
        if (active) {
            unsigned rej = 0;
            const float r = abk::clip_combine_sorted_mem(col, stride, n, 0, n, 0, nanf(""), false,
                                                         0.0f, rej, sigma_low, sigma_high,
                                                         max_iter);
            __stcs(out + p, r);
            local_rej += rej;
        }
    }
    unsigned w = __reduce_add_sync(0xffffffffu, (unsigned)local_rej);
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(rejected_total, (unsigned long long)w);
}

ab_status launch_deep(ab_ctx *ctx, const FrameTable &ft, int n_frames, size_t npix, float sl,
                      float sh, unsigned max_iter, float *d_out, unsigned long long *d_rej,
                      cudaStream_t stream) {
    int n_pad = 2;
    while (n_pad < n_frames) n_pad <<= 1;
    const size_t smem = (size_t)n_pad * DEEP_THREADS * sizeof(float);
    AB_CUDA(ctx, cudaFuncSetAttribute(stack_deep_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    int per_sm = 0;
    AB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stack_deep_kernel,
                                                               DEEP_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const size_t need = (npix + DEEP_THREADS - 1) / DEEP_THREADS;
    size_t grid = (size_t)ctx->sm_count * per_sm * 4;
    if (grid > need) grid = need;
    stack_deep_kernel<<<(unsigned)grid, DEEP_THREADS, smem, stream>>>(
        ft, n_frames, n_pad, npix, sl, sh, max_iter, d_out, d_rej);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

}  // namespace

ab_status ab_stack_launch(ab_ctx *ctx, const FrameTable &ft, int n_frames, size_t npix,
                          float sigma_low, float sigma_high, unsigned max_iter, float *d_out,
                          unsigned long long *d_rejected, cudaStream_t stream) {
    AB_CUDA(ctx, cudaMemsetAsync(d_rejected, 0, sizeof(unsigned long long), stream));
    if (npix == 0) return AB_OK;
    if (n_frames > AB_MAX_FRAMES)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "n_frames %d > AB_MAX_FRAMES %d", n_frames,
                       AB_MAX_FRAMES);
#define AB_CASE(P)                                                                              \
    case P:                                                                                     \
        return abk::launch_tile<P>(ctx, ft, npix, sigma_low, sigma_high, max_iter, d_out,      \
                                   d_rejected, stream);
    switch (n_frames) {
        AB_CASE(1) AB_CASE(2) AB_CASE(3) AB_CASE(4) AB_CASE(5) AB_CASE(6) AB_CASE(7) AB_CASE(8)
        AB_CASE(9) AB_CASE(10) AB_CASE(11) AB_CASE(12) AB_CASE(13) AB_CASE(14) AB_CASE(15)
        AB_CASE(16) AB_CASE(17) AB_CASE(18) AB_CASE(19) AB_CASE(20) AB_CASE(21) AB_CASE(22)
        AB_CASE(23) AB_CASE(24) AB_CASE(25) AB_CASE(26) AB_CASE(27) AB_CASE(28) AB_CASE(29)
        AB_CASE(30) AB_CASE(31) AB_CASE(32) AB_CASE(36) AB_CASE(40) AB_CASE(48) AB_CASE(56)
        AB_CASE(64)
        default:
            break;
    }
#undef AB_CASE
    return launch_deep(ctx, ft, n_frames, npix, sigma_low, sigma_high, max_iter, d_out,
                       d_rejected, stream);
}
