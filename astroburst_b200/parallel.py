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
from ._lib import ALLREDUCE_FN, Context, ImageStatsC, default_context, lib

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


def allreduce_rejected(rejected_tensor, group=None):
    """StackResult.rejected_pixels of a row-band sharded stack = sum over ranks (in place)."""
    import torch.distributed as dist

    dist.all_reduce(rejected_tensor, group=group)
    return rejected_tensor
