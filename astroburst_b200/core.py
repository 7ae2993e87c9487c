"""Host-side mirror of the reference's `core::stacking` functions over the C ABI.

Names, argument meaning and error behaviour follow src-tauri/src/core/stacking/combine.rs
so the parity tests read like the reference's own tests.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import AstroBurstError, Context, StackConfigC, default_context, lib


@dataclass
class StackConfig:
    """types/stacking.rs:4-20 (defaults identical)."""
    sigma_low: float = 3.0
    sigma_high: float = 3.0
    max_iterations: int = 5
    align: bool = True

    def to_c(self) -> StackConfigC:
        return StackConfigC(self.sigma_low, self.sigma_high, int(self.max_iterations),
                            1 if self.align else 0)


@dataclass
class StackResult:
    """types/stacking.rs:22-28."""
    image: np.ndarray
    frame_count: int
    rejected_pixels: int
    offsets: List[Tuple[int, int]] = field(default_factory=list)


def sigma_clip_combine(values, sigma_low: float, sigma_high: float, max_iter: int,
                       ctx: Context | None = None) -> Tuple[np.float32, int]:
    """combine.rs:14-92 for one pixel's sample list, evaluated on the device."""
    ctx = ctx or default_context()
    v = np.ascontiguousarray(values, dtype=np.float32).reshape(-1)
    mean, rej = C.c_float(0.0), C.c_uint32(0)
    ctx.check(lib().ab_sigma_clip_combine(ctx.handle, v.ctypes.data, v.size, sigma_low,
                                          sigma_high, int(max_iter), C.byref(mean),
                                          C.byref(rej)))
    return np.float32(mean.value), int(rej.value)


def stack_images(images: Sequence[np.ndarray], config: StackConfig,
                 ctx: Context | None = None, out: np.ndarray | None = None) -> StackResult:
    """combine.rs:94-193.  Host arrays in, host result out (pinned arrays run at PCIe speed).

    Frames are cropped to the common minimum size like combine.rs:104-113.  `align=True`
    phase-correlates every frame against frame 0 and shifts it with the bicubic sampler first
    (combine.rs:118-147); `offsets` are the rounded (dy, dx) of each frame.
    """
    ctx = ctx or default_context()
    if len(images) == 0:
        raise AstroBurstError(_lib.AB_ERR_INVALID, "No images to stack")
    min_rows = min(int(im.shape[0]) for im in images)
    min_cols = min(int(im.shape[1]) for im in images)
    frames = []
    for im in images:
        a = im if im.shape == (min_rows, min_cols) else im[:min_rows, :min_cols]
        frames.append(np.ascontiguousarray(a, dtype=np.float32))
    if out is None:
        out = np.empty((min_rows, min_cols), np.float32)
    assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (min_rows, min_cols)
    ptrs = (C.c_void_p * len(frames))(*[f.ctypes.data for f in frames])
    cfg = config.to_c()
    rej = C.c_uint64(0)
    offs = np.zeros((len(frames), 2), np.int32)
    ctx.check(lib().ab_stack_images(ctx.handle, ptrs, len(frames), min_rows, min_cols,
                                    C.byref(cfg), out.ctypes.data, C.byref(rej), offs.ctypes.data))
    return StackResult(out, len(frames), int(rej.value), [(int(a), int(b)) for a, b in offs])


def stack_fits_units(units: Sequence[np.ndarray], bitpix, bscale, bzero, rows: int, cols: int,
                     config: StackConfig, ctx: Context | None = None, out: np.ndarray | None = None) -> StackResult:
    """stack_images on FITS data units as read from the files (load_fits_image -> decode_pixels,
    infra/fits/reader.rs:37-101, then combine.rs:94-193, align=False).  units[f]: uint8 array of
    rows*cols big-endian samples of BITPIX bitpix[f]; bscale / bzero per frame (scalars broadcast).
    The planes cross PCIe in the file's sample type and are decoded on the device."""
    ctx = ctx or default_context()
    n = len(units)
    if n == 0:
        raise AstroBurstError(_lib.AB_ERR_INVALID, "No images to stack")
    bp = np.ascontiguousarray(np.broadcast_to(np.asarray(bitpix, np.int32), (n,)))
    bs = np.ascontiguousarray(np.broadcast_to(np.asarray(bscale, np.float64), (n,)))
    bz = np.ascontiguousarray(np.broadcast_to(np.asarray(bzero, np.float64), (n,)))
    raw = []
    for u, b in zip(units, bp):
        a = np.ascontiguousarray(u).view(np.uint8).reshape(-1)
        if a.size != rows * cols * (abs(int(b)) // 8):
            raise AstroBurstError(_lib.AB_ERR_INVALID, f"data unit of {a.size} bytes for {rows}x{cols} BITPIX {b}")
        raw.append(a)
    if out is None:
        out = np.empty((rows, cols), np.float32)
    assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (rows, cols)
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in raw])
    cfg = config.to_c()
    rej = C.c_uint64(0)
    ctx.check(lib().ab_stack_fits_units(ctx.handle, ptrs, bp.ctypes.data, bs.ctypes.data, bz.ctypes.data, n, rows,
                                        cols, C.byref(cfg), out.ctypes.data, C.byref(rej)))
    return StackResult(out, n, int(rej.value), [(0, 0)] * n)


def stack_images_device(frame_ptrs: Sequence[int], npix: int, config: StackConfig, out_ptr: int,
                        rejected_ptr: int, stream: int = 0, ctx: Context | None = None) -> None:
    """Device-resident stack: raw device pointers (e.g. torch `data_ptr()`), enqueued on
    `stream` (a cudaStream_t handle; 0 = the CUDA legacy default stream) without synchronizing."""
    ctx = ctx or default_context()
    if len(frame_ptrs) == 0:
        raise AstroBurstError(_lib.AB_ERR_INVALID, "No images to stack")
    ptrs = (C.c_void_p * len(frame_ptrs))(*[int(p) for p in frame_ptrs])
    cfg = config.to_c()
    ctx.check(lib().ab_stack_sigma_clip_dev(ctx.handle, ptrs, len(frame_ptrs), int(npix),
                                            C.byref(cfg), int(out_ptr), int(rejected_ptr),
                                            int(stream) if stream else None))
