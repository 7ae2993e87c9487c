"""Host-side mirror of the reference's `core::stacking::calibration` functions over the C ABI.

Names, argument meaning and error behaviour follow src-tauri/src/core/stacking/calibration.rs;
FITS loading (`load_fits_image`) stays with the caller, so the `create_master_*` functions take
the already-loaded frames instead of paths.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import AstroBurstError, CalibrationConfigC, Context, default_context, lib


@dataclass
class CalibrationConfig:
    """calibration.rs:8-13."""
    master_bias: Optional[np.ndarray] = None
    master_dark: Optional[np.ndarray] = None
    master_flat: Optional[np.ndarray] = None
    dark_exposure_ratio: float = 1.0


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: Optional[np.ndarray]):
    return a.ctypes.data if a is not None else None


def _check_dims(frames, what: str):
    if not frames:
        raise AstroBurstError(_lib.AB_ERR_INVALID, f"No {what} frames provided")
    shape = frames[0].shape
    for f in frames[1:]:
        if f.shape != shape:  # calibration.rs:140-145
            raise AstroBurstError(
                _lib.AB_ERR_INVALID,
                f"Dimension mismatch: expected ({shape[0]}, {shape[1]}), got {f.shape}")
    return shape


def median_combine(frames: Sequence[np.ndarray], bias: Optional[np.ndarray] = None,
                   dark: Optional[np.ndarray] = None, dark_ratio: float = 1.0,
                   ctx: Optional[Context] = None, what: str = "calibration") -> np.ndarray:
    """median_combine_row_major (calibration.rs:85-123) of `frame - bias - dark * dark_ratio`."""
    ctx = ctx or default_context()
    fr = [_f32(f) for f in frames]
    rows, cols = _check_dims(fr, what)
    b = _f32(bias) if bias is not None else None
    d = _f32(dark) if dark is not None else None
    ptrs = (C.c_void_p * len(fr))(*[f.ctypes.data for f in fr])
    out = np.empty((rows, cols), np.float32)
    ctx.check(lib().ab_median_combine(ctx.handle, ptrs, len(fr), rows, cols, _ptr(b), _ptr(d),
                                      C.c_float(dark_ratio), out.ctypes.data))
    return out


def create_master_bias(bias_frames, ctx=None) -> np.ndarray:
    """calibration.rs:125-152."""
    return median_combine(bias_frames, ctx=ctx, what="bias")


def create_master_dark(dark_frames, master_bias=None, ctx=None) -> np.ndarray:
    """calibration.rs:154-190."""
    return median_combine(dark_frames, bias=master_bias, ctx=ctx, what="dark")


def normalize_flat(flat: np.ndarray, ctx=None) -> np.ndarray:
    """calibration.rs:231-252 (returns a new array)."""
    ctx = ctx or default_context()
    out = _f32(flat).copy()
    ctx.check(lib().ab_normalize_flat(ctx.handle, out.ctypes.data, out.size))
    return out


def create_master_flat(flat_frames, master_bias=None, master_dark=None, ctx=None) -> np.ndarray:
    """calibration.rs:192-256: median of `frame - bias - dark * 1.0`, normalised to mean 1."""
    med = median_combine(flat_frames, bias=master_bias, dark=master_dark, dark_ratio=1.0, ctx=ctx,
                         what="flat")
    return normalize_flat(med, ctx=ctx)


def _calibrate(raw, bias, dark, flat, ratio, clamp, ctx):
    ctx = ctx or default_context()
    r = _f32(raw)
    b = _f32(bias) if bias is not None else None
    d = _f32(dark) if dark is not None else None
    f = _f32(flat) if flat is not None else None
    cfg = CalibrationConfigC(_ptr(b), _ptr(d), _ptr(f), float(ratio), 1 if clamp else 0)
    out = np.empty_like(r)
    ctx.check(lib().ab_calibrate_image(ctx.handle, r.ctypes.data, r.size, C.byref(cfg),
                                       out.ctypes.data))
    return out


def calibrate_image(raw: np.ndarray, config: CalibrationConfig, ctx=None) -> np.ndarray:
    """calibration.rs:47-83."""
    return _calibrate(raw, config.master_bias, config.master_dark, config.master_flat,
                      config.dark_exposure_ratio, True, ctx)


def subtract_bias(image, master_bias, ctx=None) -> np.ndarray:
    """calibration.rs:15-17."""
    return _calibrate(image, master_bias, None, None, 1.0, False, ctx)


def subtract_dark(image, master_dark, exposure_ratio: float, ctx=None) -> np.ndarray:
    """calibration.rs:19-25."""
    return _calibrate(image, None, master_dark, None, exposure_ratio, False, ctx)


def divide_flat(image, master_flat, ctx=None) -> np.ndarray:
    """calibration.rs:27-45."""
    return _calibrate(image, None, None, master_flat, 1.0, False, ctx)


def sigma_clipped_mean_stack(frames: Sequence[np.ndarray], sigma_low: float = 2.5, sigma_high: float = 3.0,
                             max_iterations: int = 5, ctx: Optional[Context] = None):
    """sigma_clipped_mean_stack of the batch pipeline (core/imaging/calibration_pipeline.rs:317-378;
    BatchStackConfig defaults :27-36) -> (stacked image, per-frame rejection counts).  Every sample
    takes part (NaN ordered last), MAD-based clip in every iteration, f32 mean in frame order."""
    ctx = ctx or default_context()
    fr = [_f32(f) for f in frames]
    if not fr:
        raise AstroBurstError(_lib.AB_ERR_INVALID, "no frames")
    rows, cols = fr[0].shape
    for f in fr[1:]:
        if f.shape != (rows, cols):  # run_batch_pipeline, calibration_pipeline.rs:131-141
            raise AstroBurstError(_lib.AB_ERR_INVALID, "All frames must match")
    ptrs = (C.c_void_p * len(fr))(*[f.ctypes.data for f in fr])
    out = np.empty((rows, cols), np.float32)
    counts = np.zeros(len(fr), np.uint64)
    ctx.check(lib().ab_sigma_clipped_mean_stack(ctx.handle, ptrs, len(fr), rows, cols, C.c_float(sigma_low),
                                                C.c_float(sigma_high), int(max_iterations), out.ctypes.data,
                                                counts.ctypes.data))
    return out, [int(c) for c in counts]
