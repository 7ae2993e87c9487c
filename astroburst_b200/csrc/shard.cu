// shard.cu -- frame-sharded stacks over the GPUs of one box through PEER MEMORY (NVLink), behind
// the C ABI (BASELINE config 5: 256 x 4096^2, 32 frames per GPU).
//
// Iteration 0 of sigma_clip_combine (combine.rs:37-48) needs every sample of a pixel in one place:
// its median / MAD do not decompose over frame shards.  So rank r combines row band r
// of ALL frames; the frames it does not own are read where they are:
//   * every rank allocates its frames as one [n_local, rows, cols] cube, exports it with CUDA IPC
//     (ab_ipc_export) and maps the other ranks' cubes (ab_ipc_open);
//   * ab_stack_frame_sharded_dev launches the ordinary stack kernel on a frame table that points
//     into the owners' cubes -- the kernel's 128-byte line fetches go over NVLink for 7/8 of the
//     frames, overlapped tile by tile with the sorting of the previous tile: transfer and compute
//     are ONE kernel, there is no staging copy and no NCCL call on the data path;
//   * StackResult.rejected_pixels (the path's only reduction) is folded into the same stream: a
//     one-warp epilogue adds this rank's count to every rank's counter with system-scope atomics
//     on the peer mappings, so after the step's barrier each rank holds the global sum.
// ab_reshard_frames_to_bands_dev is the staged alternative (pull copies of the foreign bands).
// Cross-rank ordering (barrier before reading counters / before freeing cubes) is the host's job.
#include <cstring>

#include "ab_internal.h"

namespace {

struct RejTargets {
    unsigned long long *p[AB_MAX_RANKS];
    int n;
};

__global__ void rej_fanout_kernel(const unsigned long long *__restrict__ src, RejTargets t) {
    const int i = threadIdx.x;
    if (i < t.n) atomicAdd_system(t.p[i], *src);
}

// band r of `rows`: contiguous, sizes differ by at most one row (parallel.row_band)
void row_band(size_t rows, int rank, int world, size_t *r0, size_t *r1) {
    const size_t base = rows / (size_t)world, extra = rows % (size_t)world;
    const size_t lo = (size_t)rank * base + ((size_t)rank < extra ? (size_t)rank : extra);
    *r0 = lo;
    *r1 = lo + base + ((size_t)rank < extra ? 1 : 0);
}
int frame_owner(size_t f, size_t n_frames, int world) {
    return n_frames >= (size_t)world ? (int)(f * (size_t)world / n_frames) : (int)(f % (size_t)world);
}

ab_status check_shard_args(ab_ctx *ctx, const float *const *cube_bases, int world, int rank,
                           size_t n_frames) {
    if (!ctx) return AB_ERR_INVALID;
    if (!cube_bases || world < 1 || world > AB_MAX_RANKS || rank < 0 || rank >= world)
        return ab_fail(ctx, AB_ERR_INVALID, "bad world/rank (%d/%d)", rank, world);
    if (n_frames == 0) return ab_fail(ctx, AB_ERR_INVALID, "No images to stack");
    if (n_frames > AB_MAX_FRAMES)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "n_frames %zu > AB_MAX_FRAMES %d", n_frames,
                       AB_MAX_FRAMES);
    return AB_OK;
}

// pointer to row r0 of frame f inside its owner's cube (frames of one owner are consecutive)
ab_status band_pointers(ab_ctx *ctx, const float *const *cube_bases, int world, size_t n_frames,
                        size_t rows, size_t cols, size_t r0, const float **out) {
    size_t counts[AB_MAX_RANKS] = {};
    for (size_t f = 0; f < n_frames; ++f) {
        const int o = frame_owner(f, n_frames, world);
        if (!cube_bases[o]) return ab_fail(ctx, AB_ERR_INVALID, "null cube base of rank %d", o);
        out[f] = cube_bases[o] + (counts[o] * rows + r0) * cols;
        ++counts[o];
    }
    return AB_OK;
}

// ---------------------------------------------------------------------------- peer collectives
// Cross-GPU barrier over the mapped d_small blocks: rank r stores the epoch into flag[r] of EVERY
// rank (release, system scope), then waits until its own flag[q] has reached the epoch for every q
// (acquire).  A bounded spin: a missing peer sets the error word instead of hanging the GPU.
struct PeerArgs {
    char *small[AB_MAX_RANKS];
    int world, rank;
};
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__global__ void peer_barrier_kernel(PeerArgs pa, unsigned long long epoch) {
    const int q = threadIdx.x;
    if (q >= pa.world) return;
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned long long *>(pa.small[q] + AB_SMALL_PEER_FLAGS) + pa.rank, epoch);
    const unsigned long long *mine =
        reinterpret_cast<const unsigned long long *>(pa.small[pa.rank] + AB_SMALL_PEER_FLAGS) + q;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) < epoch) {
        if (clock64() - t0 > 4000000000ll) {  // ~2 s: a peer never arrived
            *reinterpret_cast<unsigned *>(pa.small[pa.rank] + AB_SMALL_PEER_FLAGS + AB_MAX_RANKS * 8) = 1u;
            break;
        }
        __nanosleep(200);
    }
}
// dst[i] = sum over ranks (in rank order) of src[i]; every rank computes the whole result itself
__global__ void __launch_bounds__(1024)
peer_sum_u32_kernel(PeerArgs pa, size_t src_off, size_t dst_off, size_t count) {
    unsigned *dst = reinterpret_cast<unsigned *>(pa.small[pa.rank] + dst_off);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count / 4; i += stride) {
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (int q = 0; q < pa.world; ++q) {
            const uint4 v = __ldcv(reinterpret_cast<const uint4 *>(pa.small[q] + src_off) + i);
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
        }
        reinterpret_cast<uint4 *>(dst)[i] = acc;
    }
}
__global__ void peer_scalars_kernel(PeerArgs pa, PeerScalarOp op) {
    const int k = threadIdx.x;
    if (k >= op.n) return;
    char *me = pa.small[pa.rank];
    if (op.kind[k] == 1) {
        unsigned long long acc = 0;
        for (int q = 0; q < pa.world; ++q)
            acc += __ldcv(reinterpret_cast<const unsigned long long *>(pa.small[q] + op.src_off[k]));
        *reinterpret_cast<unsigned long long *>(me + op.dst_off[k]) = acc;
    } else {
        double acc = __ldcv(reinterpret_cast<const double *>(pa.small[0] + op.src_off[k]));
        for (int q = 1; q < pa.world; ++q) {
            const double v = __ldcv(reinterpret_cast<const double *>(pa.small[q] + op.src_off[k]));
            acc = op.kind[k] == 2 ? fmax(acc, v) : acc + v;
        }
        *reinterpret_cast<double *>(me + op.dst_off[k]) = acc;
    }
}
PeerArgs peer_args(ab_ctx *ctx) {
    PeerArgs pa{};
    pa.world = ctx->peer_world;
    pa.rank = ctx->peer_rank;
    for (int q = 0; q < ctx->peer_world; ++q) pa.small[q] = ctx->peer_small[q];
    return pa;
}
ab_status peer_barrier(ab_ctx *ctx, cudaStream_t s) {
    peer_barrier_kernel<<<1, 32, 0, s>>>(peer_args(ctx), ++ctx->peer_epoch);
    AB_LAUNCH_CHECK(ctx);
    return AB_OK;
}

}  // namespace

ab_status ab_peer_allreduce(ab_ctx *ctx, size_t src_off, size_t dst_off, size_t count,
                            const PeerScalarOp &op, cudaStream_t s) {
    if (ctx->peer_world < 1) return ab_fail(ctx, AB_ERR_INVALID, "no peer group (ab_peer_connect)");
    if ((count & 3) || (src_off & 15) || (dst_off & 15)) return ab_fail(ctx, AB_ERR_INVALID, "peer all-reduce alignment");
    ab_status st = peer_barrier(ctx, s);  // every rank's sources are complete
    if (st != AB_OK) return st;
    if (count) {
        peer_sum_u32_kernel<<<32, 1024, 0, s>>>(peer_args(ctx), src_off, dst_off, count);
        AB_LAUNCH_CHECK(ctx);
    }
    if (op.n) {
        peer_scalars_kernel<<<1, 32, 0, s>>>(peer_args(ctx), op);
        AB_LAUNCH_CHECK(ctx);
    }
    return peer_barrier(ctx, s);  // every rank has read every source: they may be overwritten
}
ab_status ab_peer_check(ab_ctx *ctx, cudaStream_t s) {
    if (ctx->peer_world < 1) return AB_OK;
    unsigned err = 0;
    // on the caller's stream: a synchronous cudaMemcpy is a NULL-stream command, which would order
    // this thread behind kernels other host threads have in the NULL stream (a spinning barrier
    // of another context's peer group, say) -- a cross-rank deadlock with several contexts per GPU
    AB_CUDA(ctx, cudaMemcpyAsync(&err, ctx->peer_small[ctx->peer_rank] + AB_SMALL_PEER_FLAGS + AB_MAX_RANKS * 8, 4,
                                 cudaMemcpyDeviceToHost, s));
    AB_CUDA(ctx, cudaStreamSynchronize(s));
    if (err) return ab_fail(ctx, AB_ERR_CUDA, "peer barrier timed out (a rank of the peer group did not arrive)");
    return AB_OK;
}

extern "C" {

// ---- peer group: every rank exports its context's scratch block, all ranks map all blocks
ab_status ab_peer_export(ab_ctx *ctx, void *handle64_out) {
    if (!ctx || !handle64_out) return AB_ERR_INVALID;
    return ab_ipc_export(ctx, ctx->d_small, handle64_out);
}
ab_status ab_peer_connect(ab_ctx *ctx, const void *handles64, int world, int rank) {
    if (!ctx || !handles64 || world < 1 || world > AB_MAX_RANKS || rank < 0 || rank >= world)
        return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    if (ctx->peer_world) return ab_fail(ctx, AB_ERR_INVALID, "peer group already connected");
    DeviceGuard g(ctx->device);
    for (int q = 0; q < world; ++q) {
        if (q == rank) {
            ctx->peer_small[q] = static_cast<char *>(ctx->d_small);
            continue;
        }
        void *p = nullptr;
        ab_status st = ab_ipc_open(ctx, static_cast<const char *>(handles64) + 64 * q, &p);
        if (st != AB_OK) {
            for (int k = 0; k < q; ++k)
                if (k != rank && ctx->peer_small[k]) cudaIpcCloseMemHandle(ctx->peer_small[k]);
            return st;
        }
        ctx->peer_small[q] = static_cast<char *>(p);
    }
    ctx->peer_world = world;
    ctx->peer_rank = rank;
    ctx->peer_epoch = 0;
    return AB_OK;
}
ab_status ab_peer_disconnect(ab_ctx *ctx) {
    if (!ctx) return AB_ERR_INVALID;
    CtxLock lk(ctx->mu);
    if (!ctx->peer_world) return AB_OK;
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    for (int q = 0; q < ctx->peer_world; ++q)
        if (q != ctx->peer_rank && ctx->peer_small[q]) cudaIpcCloseMemHandle(ctx->peer_small[q]);
    for (int q = 0; q < AB_MAX_RANKS; ++q) ctx->peer_small[q] = nullptr;
    ctx->peer_world = 0;
    return AB_OK;
}

void ab_row_band(size_t rows, int rank, int world, size_t *r0, size_t *r1) {
    size_t a = 0, b = 0;
    if (world > 0 && rank >= 0 && rank < world) row_band(rows, rank, world, &a, &b);
    if (r0) *r0 = a;
    if (r1) *r1 = b;
}

int ab_frame_owner(size_t f, size_t n_frames, int world) {
    if (world < 1 || n_frames == 0) return -1;
    return frame_owner(f, n_frames, world);
}

ab_status ab_ipc_export(ab_ctx *ctx, const void *dev_ptr, void *handle64_out) {
    if (!ctx || !dev_ptr || !handle64_out) return AB_ERR_INVALID;
    DeviceGuard g(ctx->device);
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    AB_CUDA(ctx, cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
    std::memcpy(handle64_out, &h, sizeof(h));
    return AB_OK;
}

ab_status ab_stack_frame_sharded_dev(ab_ctx *ctx, const float *const *cube_bases, int world,
                                     int rank, size_t n_frames, size_t rows, size_t cols,
                                     const ab_stack_config *cfg, float *d_out_band,
                                     uint64_t *const *d_rejected_targets, int n_targets,
                                     void *stream) {
    ab_status st = check_shard_args(ctx, cube_bases, world, rank, n_frames);
    if (st != AB_OK) return st;
    if (!cfg) return ab_fail(ctx, AB_ERR_INVALID, "null config");
    if (cfg->align)
        return ab_fail(ctx, AB_ERR_UNSUPPORTED, "align=true needs whole frames: use ab_stack_images");
    if (n_targets < 0 || n_targets > AB_MAX_RANKS || (n_targets && !d_rejected_targets))
        return ab_fail(ctx, AB_ERR_INVALID, "bad rejected targets");
    size_t r0, r1;
    row_band(rows, rank, world, &r0, &r1);
    const size_t npix = (r1 - r0) * cols;
    if (npix && !d_out_band) return ab_fail(ctx, AB_ERR_INVALID, "null output band");
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    FrameTable ft;
    for (size_t f = 0; f < AB_MAX_FRAMES; ++f) ft.p[f] = nullptr;
    if ((st = band_pointers(ctx, cube_bases, world, n_frames, rows, cols, r0, ft.p)) != AB_OK)
        return st;
    cudaStream_t s = (cudaStream_t)stream;
    // this rank's count lands in a scratch word (ring: calls in flight on different streams)
    unsigned long long *d_local = reinterpret_cast<unsigned long long *>(
                                      static_cast<char *>(ctx->d_small) + (776 << 10)) +
                                  (ctx->work_seq.fetch_add(1) % 64);
    st = ab_stack_launch(ctx, ft, (int)n_frames, npix, cfg->sigma_low, cfg->sigma_high,
                         cfg->max_iterations, d_out_band, d_local, s, world > 1);
    if (st != AB_OK) return st;
    if (n_targets) {
        RejTargets t{};
        t.n = n_targets;
        for (int i = 0; i < n_targets; ++i) {
            if (!d_rejected_targets[i]) return ab_fail(ctx, AB_ERR_INVALID, "null rejected target %d", i);
            t.p[i] = reinterpret_cast<unsigned long long *>(d_rejected_targets[i]);
        }
        rej_fanout_kernel<<<1, 32, 0, s>>>(d_local, t);
        AB_LAUNCH_CHECK(ctx);
    }
    return AB_OK;
}

ab_status ab_reshard_frames_to_bands_dev(ab_ctx *ctx, const float *const *cube_bases, int world,
                                         int rank, size_t n_frames, size_t rows, size_t cols,
                                         float *d_band_cube, void *stream) {
    ab_status st = check_shard_args(ctx, cube_bases, world, rank, n_frames);
    if (st != AB_OK) return st;
    size_t r0, r1;
    row_band(rows, rank, world, &r0, &r1);
    const size_t band_px = (r1 - r0) * cols;
    if (band_px == 0) return AB_OK;
    if (!d_band_cube) return ab_fail(ctx, AB_ERR_INVALID, "null band cube");
    CtxLock lk(ctx->mu);
    DeviceGuard g(ctx->device);
    const float *src[AB_MAX_FRAMES];
    if ((st = band_pointers(ctx, cube_bases, world, n_frames, rows, cols, r0, src)) != AB_OK)
        return st;
    cudaStream_t s = (cudaStream_t)stream;
    // consecutive frames of one owner are one strided 2-D block: [frames] x [band bytes]
    size_t f = 0;
    while (f < n_frames) {
        const int o = frame_owner(f, n_frames, world);
        size_t e = f + 1;
        while (e < n_frames && frame_owner(e, n_frames, world) == o) ++e;
        AB_CUDA(ctx, cudaMemcpy2DAsync(d_band_cube + f * band_px, band_px * sizeof(float), src[f],
                                       rows * cols * sizeof(float), band_px * sizeof(float), e - f,
                                       cudaMemcpyDeviceToDevice, s));
        f = e;
    }
    return AB_OK;
}

}  // extern "C"
