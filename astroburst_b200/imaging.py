"""Host-side mirror of the reference's `core::imaging`, `core::compose` and resampling functions
over the C ABI (host numpy buffers in, host results out).  Names and argument meaning follow the
Rust functions they replace (file:line cited per function, paths relative to src-tauri/src/)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import (AffineC, AstroBurstError, AutoStfConfigC, BlendWeightC, Context, ImageStatsC,
                   LevelsParamsC, ScnrConfigC, StfParamsC, default_context, lib)

ImageStats = ImageStatsC      # types/image.rs:2-10
StfParams = StfParamsC        # types/image.rs:35-40


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


# ------------------------------------------------------------------ stats.rs
def compute_image_stats(data, ctx: Context | None = None) -> ImageStats:
    """stats.rs:15-23 (exact <= 4,000,000 px, histogram algorithm above)."""
    return _stats(data, 0, 0.0, 0.0, ctx)


def compute_image_stats_exact(data, ctx=None) -> ImageStats:
    return _stats(data, 1, 0.0, 0.0, ctx)


def compute_image_stats_hist(data, ctx=None) -> ImageStats:
    return _stats(data, 2, 0.0, 0.0, ctx)


def compute_image_stats_with_known_range(data, known_min, known_max, ctx=None) -> ImageStats:
    """stats.rs:25-41."""
    return _stats(data, 3, known_min, known_max, ctx)


def _stats(data, mode, kmin, kmax, ctx):
    ctx = ctx or default_context()
    d = _f32(data).reshape(-1)
    out = ImageStatsC()
    ctx.check(lib().ab_compute_image_stats(ctx.handle, d.ctypes.data, d.size, mode, kmin, kmax,
                                           C.addressof(out)))
    return out


def build_histogram(data, n_bins, dmin, dmax, ctx=None):
    """stats.rs:378-421 -> (bins u32[n_bins], bin_edges f64[n_bins+1])."""
    ctx = ctx or default_context()
    d = _f32(data).reshape(-1)
    bins = np.zeros(n_bins, np.uint32)
    edges = np.zeros(n_bins + 1, np.float64)
    ctx.check(lib().ab_histogram_u32(ctx.handle, d.ctypes.data, d.size, n_bins, dmin, dmax,
                                     bins.ctypes.data, edges.ctypes.data))
    return bins, edges


def compute_histogram_with_stats(data, stats: ImageStats, ctx=None):
    """stats.rs:373-376 (HISTOGRAM_BINS = 65536 over [stats.min, stats.max])."""
    return build_histogram(data, 65536, stats.min, stats.max, ctx)


def downsample_histogram(bins, target_bins):
    b = np.ascontiguousarray(bins, np.uint32)
    out = np.zeros(min(target_bins, b.size), np.uint32)
    st = lib().ab_downsample_histogram(b.ctypes.data, b.size, target_bins, out.ctypes.data)
    if st:
        raise AstroBurstError(st, "ab_downsample_histogram")
    return out


def auto_stf(stats: ImageStats, target_bg=0.25, shadow_k=-2.8) -> StfParams:
    """stf.rs:13-39 with AutoStfConfig defaults (types/image.rs:57-64)."""
    cfg = AutoStfConfigC(target_bg, shadow_k)
    out = StfParamsC()
    st = lib().ab_auto_stf(C.addressof(stats), C.addressof(cfg), C.addressof(out))
    if st:
        raise AstroBurstError(st, "ab_auto_stf")
    return out


def linked_stats(r: ImageStats, g: ImageStats, b: ImageStats) -> ImageStats:
    out = ImageStatsC()
    lib().ab_linked_stats(C.addressof(r), C.addressof(g), C.addressof(b), C.addressof(out))
    return out


def select_wb_reference(r: ImageStats, g: ImageStats, b: ImageStats):
    out = (C.c_double * 3)()
    lib().ab_select_wb_reference(C.addressof(r), C.addressof(g), C.addressof(b), out)
    return tuple(out)


# ------------------------------------------------------------------ stf.rs / stretch.rs
def apply_stf(data, params: StfParams, stats: ImageStats, ctx=None) -> np.ndarray:
    """stf.rs:89-102 -> u8."""
    ctx = ctx or default_context()
    d = _f32(data)
    out = np.empty(d.shape, np.uint8)
    ctx.check(lib().ab_apply_stf_u8(ctx.handle, d.ctypes.data, d.size, C.addressof(params),
                                    C.addressof(stats), out.ctypes.data))
    return out


def apply_stf_f32(data, params: StfParams, stats: ImageStats, rgb_variant=False, ctx=None):
    """stf.rs:104-120; rgb_variant=True is core/compose/rgb.rs:191-207."""
    ctx = ctx or default_context()
    d = _f32(data)
    out = np.empty(d.shape, np.float32)
    ctx.check(lib().ab_apply_stf_f32(ctx.handle, d.ctypes.data, d.size, C.addressof(params),
                                     C.addressof(stats), 1 if rgb_variant else 0, out.ctypes.data))
    return out


def arcsinh_stretch_with_stats(data, dmin, dmax, factor, gamma=1.0, ctx=None):
    """stretch.rs:10-46."""
    ctx = ctx or default_context()
    d = _f32(data)
    out = np.empty(d.shape, np.float32)
    ctx.check(lib().ab_arcsinh_stretch(ctx.handle, d.ctypes.data, d.size, dmin, dmax, factor, gamma,
                                       out.ctypes.data))
    return out


def arcsinh_stretch(data, factor, ctx=None):
    """stretch.rs:5-8."""
    st = compute_image_stats(data, ctx)
    return arcsinh_stretch_with_stats(data, np.float32(st.min), np.float32(st.max), factor, 1.0, ctx)


def arcsinh_stretch_rgb(r, g, b, factor, ctx=None):
    """stretch.rs:48-90 without explicit bounds: global min/max of the three stats (G20)."""
    sr, sg, sb = (compute_image_stats(x, ctx) for x in (r, g, b))
    gmin = min(np.float32(sr.min), np.float32(sg.min), np.float32(sb.min))
    gmax = max(np.float32(sr.max), np.float32(sg.max), np.float32(sb.max))
    return tuple(arcsinh_stretch_with_stats(x, gmin, gmax, factor, 1.0, ctx) for x in (r, g, b))


# ------------------------------------------------------------------ compose
def blend_channels(channels: Sequence[np.ndarray], weights, rows: int, cols: int, ctx=None):
    """channel_blend.rs:13-70.  weights: (channel_idx, r, g, b) tuples."""
    ctx = ctx or default_context()
    ch = [_f32(c).reshape(-1) for c in channels]
    n = len(ch)
    ptrs = (C.c_void_p * max(n, 1))(*[c.ctypes.data for c in ch])
    lens = (C.c_size_t * max(n, 1))(*[c.size for c in ch])
    w = (BlendWeightC * max(len(weights), 1))(*[BlendWeightC(int(i), r, g, b) for i, r, g, b in weights])
    r = np.empty((rows, cols), np.float32)
    g = np.empty((rows, cols), np.float32)
    b = np.empty((rows, cols), np.float32)
    ctx.check(lib().ab_blend_channels(ctx.handle, ptrs, lens, n, w, len(weights), rows, cols,
                                      r.ctypes.data, g.ctypes.data, b.ctypes.data))
    return r, g, b


def apply_scnr_inplace(r, g, b, method="average", amount=1.0, preserve_luminance=False, ctx=None):
    """scnr.rs:18-53: in place; silently returns on a shape mismatch (scnr.rs:24-26)."""
    if r.shape != g.shape or g.shape != b.shape:
        return
    ctx = ctx or default_context()
    for a in (r, g, b):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    cfg = ScnrConfigC(0 if method == "average" else 1, amount, 1 if preserve_luminance else 0)
    ctx.check(lib().ab_apply_scnr(ctx.handle, r.ctypes.data, g.ctypes.data, b.ctypes.data, r.size,
                                  C.addressof(cfg)))


def scale(data, factor, ctx=None):
    """compose/rgb.rs:127-130, cmd/compose/color.rs:21-49: v * factor."""
    ctx = ctx or default_context()
    d = _f32(data)
    out = np.empty(d.shape, np.float32)
    ctx.check(lib().ab_scale(ctx.handle, d.ctypes.data, d.size, factor, out.ctypes.data))
    return out


def apply_levels(data, black=0.0, gamma=1.0, white=1.0, ctx=None):
    """curves.rs:31-50."""
    ctx = ctx or default_context()
    d = _f32(data)
    out = np.empty(d.shape, np.float32)
    p = LevelsParamsC(black, gamma, white)
    ctx.check(lib().ab_apply_levels(ctx.handle, d.ctypes.data, d.size, C.addressof(p), out.ctypes.data))
    return out


def spline_lut(points) -> np.ndarray:
    """SplineLut::from_points, curves.rs:70-95 -> 4096 f32."""
    pts = np.ascontiguousarray(points, np.float64).reshape(-1)
    lut = np.empty(4096, np.float32)
    st = lib().ab_spline_lut(pts.ctypes.data if pts.size else None, pts.size // 2, lut.ctypes.data)
    if st:
        raise AstroBurstError(st, "ab_spline_lut")
    return lut


def apply_curve(data, lut, ctx=None):
    """curves.rs:186-198."""
    ctx = ctx or default_context()
    d = _f32(data)
    l = _f32(lut)
    assert l.size == 4096
    out = np.empty(d.shape, np.float32)
    ctx.check(lib().ab_apply_curve(ctx.handle, d.ctypes.data, d.size, l.ctypes.data, out.ctypes.data))
    return out


def apply_tone_composite(r, g, b, stf3, stats3, levels3=None, curves3=None, scnr=None, ctx=None):
    """The pixel chain of apply_tone_composite_cmd (cmd/processing/curves.rs:111-151) in one sweep:
    STF -> levels -> curves -> SCNR.  stf3 / stats3: per-channel StfParams / ImageStats;
    levels3: three (black, gamma, white) or None; curves3: three 4096-entry LUTs (spline_lut) or
    None; scnr: (method, amount, preserve_luminance) or None.  -> (r, g, b) float32 arrays."""
    ctx = ctx or default_context()
    r, g, b = _f32(r), _f32(g), _f32(b)
    if not (r.shape == g.shape == b.shape):
        raise ValueError("channel dimension mismatch")
    ch = _lib.ToneChainC()
    luts = []
    for k in range(3):
        ch.stf[k], ch.stats[k] = stf3[k], stats3[k]
        ch.levels[k] = LevelsParamsC(*(levels3[k] if levels3 is not None else (0.0, 1.0, 1.0)))
        if curves3 is not None:
            l = _f32(curves3[k])
            assert l.size == 4096
            luts.append(l)
            ch.lut[k] = l.ctypes.data
    if scnr is not None:
        method, amount, preserve = scnr
        ch.apply_scnr = 1
        ch.scnr = ScnrConfigC(0 if method == "average" else 1, amount, 1 if preserve else 0)
    outs = [np.empty(r.shape, np.float32) for _ in range(3)]
    ctx.check(lib().ab_tone_composite(ctx.handle, r.ctypes.data, g.ctypes.data, b.ctypes.data, r.size,
                                      C.addressof(ch), outs[0].ctypes.data, outs[1].ctypes.data,
                                      outs[2].ctypes.data))
    return tuple(outs)


# ------------------------------------------------------------------ bicubic gathers
def warp_image(image, transform, out_rows: int, out_cols: int, ctx=None):
    """affine.rs:663-690; transform = (a, b, tx, c, d, ty) mapping OUTPUT -> SOURCE."""
    ctx = ctx or default_context()
    d = _f32(image)
    t = AffineC(*transform)
    out = np.empty((out_rows, out_cols), np.float32)
    ctx.check(lib().ab_warp_affine_bicubic(ctx.handle, d.ctypes.data, d.shape[0], d.shape[1],
                                           C.addressof(t), out_rows, out_cols, out.ctypes.data))
    return out


def shift_image_subpixel(image, dy: float, dx: float, ctx=None):
    """stacking/align.rs:36-57."""
    ctx = ctx or default_context()
    d = _f32(image)
    out = np.empty(d.shape, np.float32)
    ctx.check(lib().ab_shift_bicubic(ctx.handle, d.ctypes.data, d.shape[0], d.shape[1], dy, dx,
                                     out.ctypes.data))
    return out


def resample_image(image, target_rows: int, target_cols: int, ctx=None):
    """imaging/resample.rs:25-61 (Err on a zero target dimension)."""
    ctx = ctx or default_context()
    d = _f32(image)
    out = np.empty((max(target_rows, 0), max(target_cols, 0)), np.float32)
    ctx.check(lib().ab_resample_bicubic(ctx.handle, d.ctypes.data, d.shape[0], d.shape[1],
                                        target_rows, target_cols, out.ctypes.data))
    return out
