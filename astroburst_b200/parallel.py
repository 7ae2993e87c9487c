"""Multi-GPU plumbing for the hot path (one process per GPU, torch.distributed over NCCL/NVLink;
gloo on CPU for the tests).  SURVEY 8(e): every kernel except the global statistics is
embarrassingly parallel over pixels, so images are cut into row bands with zero data-path traffic;
the only exchanges are the tiny reductions below (rejected-pixel counter, min/max, 65,536-bin
histograms).  Frame-sharded input must be re-sharded to row bands first (exact median/MAD needs
all samples of a pixel on one GPU)."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from . import _lib
from ._lib import ALLREDUCE_DEV_FN, ALLREDUCE_FN, Context, ImageStatsC, default_context, lib

_NP = {0: np.float64, 1: np.uint64, 2: np.float64, 3: np.float64}


def row_band(rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [r0, r1) owned by `rank`: contiguous bands, sizes differ by at most one row."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(rows, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def frame_to_band_plan(n_frames: int, rows: int, world: int):
    """All-to-all plan for frame-sharded -> row-band sharded input: entry [src][dst] lists
    (frame, r0, r1) blocks rank `src` sends to rank `dst` (its own band stays local)."""
    plan = [[[] for _ in range(world)] for _ in range(world)]
    for f in range(n_frames):
        src = f * world // n_frames if n_frames >= world else f % world
        for dst in range(world):
            r0, r1 = row_band(rows, dst, world)
            if r1 > r0:
                plan[src][dst].append((f, r0, r1))
    return plan


def make_allreduce(group=None, device=None):
    """ctypes callback for ab_image_stats_sharded_dev: reduces a host buffer over all ranks with
    torch.distributed (sum of f64 / u64 -- u64 goes through int64, counts stay < 2^63 -- min, max).
    With the NCCL backend the buffer is staged through `device`."""
    import torch
    import torch.distributed as dist

    ops = {0: dist.ReduceOp.SUM, 1: dist.ReduceOp.SUM, 2: dist.ReduceOp.MIN, 3: dist.ReduceOp.MAX}

    def _cb(user, buf, count, dtype):
        try:
            nbytes = int(count) * 8
            raw = (C.c_char * nbytes).from_address(buf)
            arr = np.frombuffer(raw, dtype=np.int64 if dtype == 1 else np.float64)
            t = torch.from_numpy(arr)
            if device is not None:
                td = t.to(device)
                dist.all_reduce(td, op=ops[dtype], group=group)
                t.copy_(td.cpu())
            else:
                dist.all_reduce(t, op=ops[dtype], group=group)
            return 0
        except Exception:  # never unwind through the C frame
            return 1

    return ALLREDUCE_FN(_cb)


def compute_image_stats_sharded(band_ptr: int, band_len: int, allreduce, known_range=None,
                                stream: int = 0, ctx: Context | None = None) -> ImageStatsC:
    """compute_image_stats (histogram algorithm, stats.rs:75-210) of an image whose row bands live
    on different GPUs.  `band_ptr` is this rank's device pointer; every rank gets the same stats."""
    ctx = ctx or default_context()
    out = ImageStatsC()
    kmin, kmax = known_range if known_range is not None else (0.0, 0.0)
    ctx.check(lib().ab_image_stats_sharded_dev(
        ctx.handle, int(band_ptr) if band_len else None, int(band_len), allreduce, None,
        1 if known_range is not None else 0, kmin, kmax, C.addressof(out),
        int(stream) if stream else None))
    return out


class _DevView:
    """__cuda_array_interface__ view of a raw device buffer (what the C callback hands over)."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False),
                                         "version": 2, "strides": None}


def make_allreduce_dev(device, group=None):
    """ctypes callback for ab_image_stats_sharded_chained_dev: one torch.distributed (NCCL)
    all-reduce of a DEVICE buffer, in place, enqueued on the caller's stream.  u64 / u32 buffers are
    summed as int64 / int32 (same bits: two's complement, counts stay far below the sign bit)."""
    import torch
    import torch.distributed as dist

    ops = {0: dist.ReduceOp.SUM, 1: dist.ReduceOp.MIN, 2: dist.ReduceOp.MAX}
    types = {0: "<f8", 1: "<i8", 2: "<i4"}

    def _cb(user, buf, count, dtype, op, stream):
        try:
            t = torch.as_tensor(_DevView(int(buf), int(count), types[dtype]), device=device)
            st = torch.cuda.ExternalStream(int(stream)) if stream else torch.cuda.default_stream(device)
            with torch.cuda.stream(st):
                dist.all_reduce(t, op=ops[op], group=group)
            return 0
        except Exception:  # never unwind through the C frame
            return 1

    return ALLREDUCE_DEV_FN(_cb)


def compute_image_stats_sharded_chained(band_ptr: int, band_len: int, allreduce_dev, known_range=None,
                                        stream: int = 0, ctx: Context | None = None) -> ImageStatsC:
    """compute_image_stats of a row-band sharded image with every reduction on the device and on
    the stream (NCCL through `allreduce_dev` = make_allreduce_dev(...)); same results as
    compute_image_stats_sharded, no host round trip between the sweeps."""
    ctx = ctx or default_context()
    out = ImageStatsC()
    kmin, kmax = known_range if known_range is not None else (0.0, 0.0)
    ctx.check(lib().ab_image_stats_sharded_chained_dev(
        ctx.handle, int(band_ptr) if band_len else None, int(band_len), allreduce_dev, None,
        1 if known_range is not None else 0, kmin, kmax, C.addressof(out),
        int(stream) if stream else None))
    return out


def connect_peer_group(ctx: Context | None = None, group=None) -> None:
    """Set up the library's peer group (shard.cu): every rank exports its context's scratch block,
    the 64-byte handles travel over torch.distributed, every rank maps all blocks.  From then on the
    sharded statistics reduce inside kernels over NVLink (compute_image_stats_peer)."""
    import torch.distributed as dist

    ctx = ctx or default_context()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    buf = (C.c_char * 64)()
    ctx.check(lib().ab_peer_export(ctx.handle, buf))
    gathered = [None] * world
    dist.all_gather_object(gathered, bytes(buf.raw), group=group)
    blob = b"".join(gathered)
    ctx.check(lib().ab_peer_connect(ctx.handle, blob, world, rank))
    dist.barrier(group)


def disconnect_peer_group(ctx: Context | None = None, group=None) -> None:
    import torch.distributed as dist

    ctx = ctx or default_context()
    ctx.synchronize()
    dist.barrier(group)  # nobody is inside a peer collective any more
    ctx.check(lib().ab_peer_disconnect(ctx.handle))
    dist.barrier(group)


def compute_image_stats_peer(band_ptr: int, band_len: int, known_range=None, stream: int = 0,
                             ctx: Context | None = None) -> ImageStatsC:
    """compute_image_stats of a row-band sharded image over the peer group (connect_peer_group):
    counters and histograms are summed by kernels that read the peers' blocks over NVLink."""
    ctx = ctx or default_context()
    out = ImageStatsC()
    kmin, kmax = known_range if known_range is not None else (0.0, 0.0)
    ctx.check(lib().ab_image_stats_peer_dev(ctx.handle, int(band_ptr) if band_len else None, int(band_len),
                                            1 if known_range is not None else 0, kmin, kmax, C.addressof(out),
                                            int(stream) if stream else None))
    return out


def allreduce_rejected(rejected_tensor, group=None):
    """StackResult.rejected_pixels of a row-band sharded stack = sum over ranks (in place)."""
    import torch.distributed as dist

    dist.all_reduce(rejected_tensor, group=group)
    return rejected_tensor


def frame_owner(f: int, n_frames: int, world: int) -> int:
    """Rank holding frame `f` of a frame-sharded stack (contiguous blocks, like the plan above)."""
    return f * world // n_frames if n_frames >= world else f % world


def reshard_frames_to_bands(local_frames, n_frames: int, rows: int, cols: int, out=None, group=None):
    """Frame-sharded -> row-band sharded input: the exchange step a
    frame-sharded deep stack needs before the exact per-pixel median/MAD.

    `local_frames`: [n_local, rows, cols] tensor holding this rank's frames in global frame order
    (frame f lives on frame_owner(f)).  Returns `out` = [n_frames, band_rows, cols]: this rank's row
    band of EVERY frame.  Zero-copy on both sides: each (frame, band) block is a contiguous slice of
    the source and of the destination cube, sent with one batched isend/irecv group (NCCL: one
    ncclGroup of ncclSend/ncclRecv over NVLink; gloo on CPU for the tests)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = [f for f in range(n_frames) if frame_owner(f, n_frames, world) == rank]
    if local_frames.shape[0] != len(mine) or tuple(local_frames.shape[1:]) != (rows, cols):
        raise ValueError(f"rank {rank} must hold {len(mine)} frames of {rows}x{cols}, "
                         f"got {tuple(local_frames.shape)}")
    if not local_frames.is_contiguous():
        raise ValueError("local_frames must be contiguous")
    r0, r1 = row_band(rows, rank, world)
    if out is None:
        out = torch.empty((n_frames, r1 - r0, cols), dtype=local_frames.dtype,
                          device=local_frames.device)
    ops = []
    local_index = {f: i for i, f in enumerate(mine)}
    for f in range(n_frames):
        src = frame_owner(f, n_frames, world)
        if src == rank:
            for dst in range(world):
                d0, d1 = row_band(rows, dst, world)
                if d1 == d0:
                    continue
                blk = local_frames[local_index[f], d0:d1]
                if dst == rank:
                    out[f].copy_(blk)
                else:
                    ops.append(dist.P2POp(dist.isend, blk, dst, group))
        elif r1 > r0:
            ops.append(dist.P2POp(dist.irecv, out[f], src, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out


def stack_frame_sharded(local_frames, n_frames: int, rows: int, cols: int, config, stream: int = 0,
                        ctx: Context | None = None, group=None):
    """Config 5: sigma-clip stack of a stack whose FRAMES are sharded over the ranks.  Re-shards to
    row bands (one all-to-all over NVLink), runs the exact local kernel on the band, all-reduces
    StackResult.rejected_pixels.  Returns (band [band_rows, cols] device tensor, rejected int64
    tensor, (r0, r1)); bit-identical to the single-GPU stack of all frames."""
    import torch
    import torch.distributed as dist

    from .core import stack_images_device

    ctx = ctx or default_context()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    cube = reshard_frames_to_bands(local_frames, n_frames, rows, cols, group=group)
    r0, r1 = row_band(rows, rank, world)
    out = torch.empty((r1 - r0, cols), dtype=torch.float32, device=cube.device)
    rej = torch.zeros(1, dtype=torch.int64, device=cube.device)
    if r1 > r0:
        stack_images_device([cube[f].data_ptr() for f in range(n_frames)], out.numel(), config,
                            out.data_ptr(), rej.data_ptr(), stream, ctx)
    allreduce_rejected(rej, group)
    return out, rej, (r0, r1)


# ---------------------------------------------------------------------------------------------------
# Peer-memory path (shard.cu): frame-sharded stacks read the owners' HBM over NVLink inside the
# stack kernel -- transfer and compute are one kernel, no staging copy, no NCCL on the data path.
class DeviceCube:
    """A cudaMalloc'd [n, rows, cols] f32 block owned by the library (ab_dev_alloc), exportable
    with CUDA IPC.  `tensor()` views it as a torch tensor (__cuda_array_interface__) for filling."""

    def __init__(self, shape, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self.shape = tuple(int(x) for x in shape)
        self.nbytes = 4 * int(np.prod(self.shape)) if len(self.shape) else 4
        p = C.c_void_p()
        self.ctx.check(lib().ab_dev_alloc(self.ctx.handle, max(self.nbytes, 1), C.byref(p)))
        self.ptr = int(p.value)
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": "<f4", "data": (self.ptr, False),
                                         "version": 2, "strides": None}

    def tensor(self):
        import torch

        return torch.as_tensor(self, device=torch.device("cuda", self.ctx.device))

    def ipc_handle(self) -> bytes:
        buf = (C.c_char * 64)()
        self.ctx.check(lib().ab_ipc_export(self.ctx.handle, self.ptr, buf))
        return bytes(buf.raw)

    def free(self):
        if self.ptr:
            self.ctx.check(lib().ab_dev_free(self.ctx.handle, self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PeerCubes:
    """Every rank's exported block, mapped into this process (ab_ipc_open under this rank's device):
    `base[q]` = rank q's block as a device pointer valid here.  `local` is a DeviceCube.  The handle
    bytes travel over torch.distributed (any host channel would do)."""

    def __init__(self, local: DeviceCube, group=None):
        import torch
        import torch.distributed as dist

        self.ctx = local.ctx
        self.local = local
        self.group = group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self.world, self.rank = world, rank
        torch.cuda.synchronize(self.ctx.device)  # the block is complete before anyone maps it
        gathered = [None] * world
        dist.all_gather_object(gathered, local.ipc_handle(), group=group)
        self._opened = []
        self.base = []
        for q in range(world):
            if q == rank:
                self.base.append(local.ptr)
                continue
            p = C.c_void_p()
            self.ctx.check(lib().ab_ipc_open(self.ctx.handle, gathered[q], C.byref(p)))
            self._opened.append(p)
            self.base.append(int(p.value))
        dist.barrier(group)

    def close(self):
        """Unmap the peers' blocks (every rank, before any owner frees its block)."""
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize(self.ctx.device)
        for p in self._opened:
            self.ctx.check(lib().ab_ipc_close(self.ctx.handle, p))
        self._opened = []
        dist.barrier(self.group)


def band_pointers_in_cubes(bases, n_frames: int, rows: int, cols: int, rank: int, world: int, elem: int = 4):
    """Pointer arithmetic of the peer-memory stack (same rule as shard.cu): frame f's band of `rank`
    starts at its owner's cube base + ((index of f among the owner's frames) * rows + r0) * cols."""
    r0, r1 = row_band(rows, rank, world)
    counts = [0] * world
    ptrs = []
    for f in range(n_frames):
        o = frame_owner(f, n_frames, world)
        ptrs.append(int(bases[o]) + ((counts[o] * rows + r0) * cols) * elem)
        counts[o] += 1
    return ptrs, (r0, r1)


def stack_frame_sharded_p2p(cubes: PeerCubes, n_frames: int, rows: int, cols: int, config, out_ptr: int,
                            rejected_targets=(), stream: int = 0):
    """Config 5 as ONE kernel per rank (ab_stack_frame_sharded_dev): this rank's row band of every
    frame is read from the owners' HBM; its reject count is added to every pointer in
    `rejected_targets` (own counter + the peers' mapped counters) by a peer-atomic epilogue.
    Returns (r0, r1).  Bit-identical to the single-GPU stack of all frames."""
    bases = (C.c_void_p * cubes.world)(*[int(b) for b in cubes.base])
    tg = (C.c_void_p * max(len(rejected_targets), 1))(*[int(t) for t in rejected_targets])
    cfg = config.to_c()
    cubes.ctx.check(lib().ab_stack_frame_sharded_dev(
        cubes.ctx.handle, bases, cubes.world, cubes.rank, int(n_frames), int(rows), int(cols), C.byref(cfg),
        int(out_ptr) if out_ptr else None, tg, len(rejected_targets), int(stream) if stream else None))
    return row_band(rows, cubes.rank, cubes.world)


def reshard_frames_to_bands_p2p(cubes: PeerCubes, n_frames: int, rows: int, cols: int, band_cube_ptr: int,
                                stream: int = 0):
    """Staged alternative behind the C ABI (ab_reshard_frames_to_bands_dev): pull this rank's band of
    every frame into a local [n_frames, band_rows, cols] cube."""
    bases = (C.c_void_p * cubes.world)(*[int(b) for b in cubes.base])
    cubes.ctx.check(lib().ab_reshard_frames_to_bands_dev(
        cubes.ctx.handle, bases, cubes.world, cubes.rank, int(n_frames), int(rows), int(cols),
        int(band_cube_ptr) if band_cube_ptr else None, int(stream) if stream else None))
    return row_band(rows, cubes.rank, cubes.world)


# ------------------------------------------------------------------------- host placement
def bind_to_gpu_numa(device_index: int) -> str:
    """Restrict this process (one per GPU) to the CPUs NVML reports as local to the GPU, so that
    pinned host buffers allocated afterwards land on the GPU's NUMA node and host<->device copies
    do not cross the socket interconnect.  Returns a one-line description of what was done; never
    raises (placement is an optimisation, results do not depend on it)."""
    import os

    if os.environ.get("AB_NO_NUMA_BIND"):
        return "off (AB_NO_NUMA_BIND)"
    try:
        import pynvml as n
        import torch

        n.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            h = n.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            h = n.nvmlDeviceGetHandleByIndex(device_index)
        words = n.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "unchanged (no local CPUs inside the current affinity mask)"
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} CPUs local to GPU {device_index} (NVML cpu affinity)"
    except Exception as e:  # no NVML, no permission, ...
        return f"unchanged ({type(e).__name__})"
