// stack.cu -- dispatcher of the sigma-clipped frame stack and the deep-stack kernel.
//
// Reference semantics: stack_images / sigma_clip_combine,
// src-tauri/src/core/stacking/combine.rs:14-92,160-182 (align=false core loop).
// The register-resident kernels (exact frame count P, see stack_kernels.cuh) are instantiated
// in stack_inst_*.cu for P = 1..32, 36, 40, 48, 56, 64; every other count <= 256 runs the
// chunked deep-stack kernel (stack_deep.cu).
#include <cfloat>
#include <cmath>

#include <cuda.h>

#include "stack_kernels.cuh"

namespace abk {

// [P, npix] f32 view of the frame cube, box = box_px pixels x P frames, no swizzle: the tile lands
// in shared memory as stage[f][pixel], exactly what the per-thread column reads expect.
bool ab_make_stack_tmap(ab_ctx *ctx, const float *base, size_t npix, size_t plane_stride_elems,
                        int P, int box_px, TmapParam *out, int n_rows) {
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
    // n_rows < P: the cube holds fewer frames than the kernel's box; the missing rows arrive as NaN
    // (out-of-bounds fill), i.e. as samples the finite filter drops
    const cuuint64_t gdim[2] = {(cuuint64_t)npix, (cuuint64_t)n_rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)plane_stride_elems * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_px, (cuuint32_t)P};
    const cuuint32_t estride[2] = {1, 1};
    CUresult r = reinterpret_cast<encode_fn>(ctx->encode_tmap_fn)(
        reinterpret_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
        const_cast<float *>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        n_rows < P ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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

ab_status ab_stack_launch(ab_ctx *ctx, const FrameTable &ft, int n_frames, size_t npix,
                          float sigma_low, float sigma_high, unsigned max_iter, float *d_out,
                          unsigned long long *d_rejected, cudaStream_t stream, bool peer_frames) {
    // (the register kernels publish the count themselves, see launch_mix; only the empty image
    // and the deep path zero it here)
    if (npix == 0) {
        AB_CUDA(ctx, cudaMemsetAsync(d_rejected, 0, sizeof(unsigned long long), stream));
        return AB_OK;
    }
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
    // 33..63 frames without a kernel of their own: the next register kernel up, its missing frames
    // delivered as NaN by the tensor map (frames the finite filter drops, combine.rs:170-175, so
    // the result is the same stack).  Needs the frames as one aligned constant-stride cube; the
    // kernel falls back to the deep path itself when the tensor map cannot be built.
    if (n_frames > 32 && n_frames < 64 && !peer_frames && npix < ((size_t)1 << 31)) {
        const ptrdiff_t stride = ft.p[1] - ft.p[0];
        bool cube = stride > 0 && (stride % 4) == 0 && (size_t)stride >= npix && ctx->stack_variant != 2;
        for (int f = 0; cube && f < n_frames; ++f)
            cube = (reinterpret_cast<uintptr_t>(ft.p[f]) & 15) == 0 && (f == 0 || ft.p[f] - ft.p[f - 1] == stride);
        if (cube) {
            const int padded = n_frames <= 36 ? 36 : n_frames <= 40 ? 40 : n_frames <= 48 ? 48 : n_frames <= 56 ? 56 : 64;
            FrameTable pt;
            for (int f = 0; f < AB_MAX_FRAMES; ++f) pt.p[f] = f < n_frames ? ft.p[f] : nullptr;  // null = padding
            switch (padded) {
                case 36: return abk::launch_tile<36>(ctx, pt, npix, sigma_low, sigma_high, max_iter, d_out, d_rejected, stream);
                case 40: return abk::launch_tile<40>(ctx, pt, npix, sigma_low, sigma_high, max_iter, d_out, d_rejected, stream);
                case 48: return abk::launch_tile<48>(ctx, pt, npix, sigma_low, sigma_high, max_iter, d_out, d_rejected, stream);
                case 56: return abk::launch_tile<56>(ctx, pt, npix, sigma_low, sigma_high, max_iter, d_out, d_rejected, stream);
                default: return abk::launch_tile<64>(ctx, pt, npix, sigma_low, sigma_high, max_iter, d_out, d_rejected, stream);
            }
        }
    }
    AB_CUDA(ctx, cudaMemsetAsync(d_rejected, 0, sizeof(unsigned long long), stream));
    DeepPre pre{};
    pre.peer_frames = peer_frames ? 1 : 0;
    return ab_stack_launch_deep(ctx, ft, n_frames, npix, sigma_low, sigma_high, max_iter, pre, d_out,
                                d_rejected, stream);
}
