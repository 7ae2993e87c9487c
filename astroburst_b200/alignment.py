"""Host-side mirror of `core::analysis::star_detection` and `core::alignment::affine` over the
C ABI (paths relative to src-tauri/src/)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List

import numpy as np

from ._lib import (AlignResultC, AstroBurstError, Context, DetectionInfoC, StarC, SubframeConfigC,
                   SubframeMetricsC, default_context, lib)

METHODS = ("affine", "rigid", "phase_correlation", "identity")  # affine.rs:101-118


@dataclass
class DetectionResult:
    """star_detection.rs:23-30."""
    stars: List[dict]
    background_median: float
    background_sigma: float
    threshold_sigma: float
    image_width: int
    image_height: int


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def detect_stars(image, sigma_threshold: float, max_stars: int = 1 << 20,
                 ctx: Context | None = None) -> DetectionResult:
    """star_detection.rs:86-258."""
    ctx = ctx or default_context()
    d = _f32(image)
    buf = (StarC * max_stars)()
    info = DetectionInfoC()
    ctx.check(lib().ab_detect_stars(ctx.handle, d.ctypes.data, d.shape[0], d.shape[1],
                                    sigma_threshold, buf, max_stars, C.addressof(info)))
    n = min(int(info.n_stars), max_stars)
    stars = [dict(x=s.x, y=s.y, flux=s.flux, fwhm=s.fwhm, eccentricity=s.eccentricity, peak=s.peak,
                  npix=int(s.npix), snr=s.snr) for s in buf[:n]]
    return DetectionResult(stars, info.background_median, info.background_sigma,
                           info.threshold_sigma, int(info.image_width), int(info.image_height))


def estimate_background(image, tile_size: int, ctx=None):
    """star_detection.rs:32-84 -> (median, sigma)."""
    ctx = ctx or default_context()
    d = _f32(image)
    med, sig = C.c_double(), C.c_double()
    ctx.check(lib().ab_estimate_background(ctx.handle, d.ctypes.data, d.shape[0], d.shape[1],
                                           tile_size, C.byref(med), C.byref(sig)))
    return med.value, sig.value


def normalize_for_detection(image, ctx=None):
    """affine.rs:24-53."""
    ctx = ctx or default_context()
    d = _f32(image)
    out = np.empty(d.shape, np.float32)
    ctx.check(lib().ab_normalize_for_detection(ctx.handle, d.ctypes.data, d.size, out.ctypes.data))
    return out


def align_channel_affine(reference, target, n_threads: int = 8, ctx=None) -> AlignResultC:
    """affine.rs:129-212 -> AffineAlignResult (transform maps OUTPUT -> SOURCE pixels)."""
    ctx = ctx or default_context()
    r, t = _f32(reference), _f32(target)
    assert r.shape == t.shape
    out = AlignResultC()
    ctx.check(lib().ab_align_affine(ctx.handle, r.ctypes.data, t.ctypes.data, r.shape[0], r.shape[1],
                                    n_threads, C.addressof(out)))
    return out


def match_stars(ref_xy, tgt_xy, max_matches: int = 256, ctx=None) -> np.ndarray:
    """build_triangles + match_triangles, affine.rs:279-384 -> [n, 4] (rx, ry, tx, ty)."""
    ctx = ctx or default_context()
    r = np.ascontiguousarray(ref_xy, np.float64).reshape(-1, 2)
    t = np.ascontiguousarray(tgt_xy, np.float64).reshape(-1, 2)
    out = np.zeros((max_matches, 4), np.float64)
    n = C.c_size_t(0)
    ctx.check(lib().ab_match_stars(ctx.handle, r.ctypes.data, r.shape[0], t.ctypes.data, t.shape[0],
                                   out.ctypes.data, max_matches, C.byref(n)))
    return out[: min(n.value, max_matches)].copy()


def ransac_affine(matches, method: str = "affine", n_threads: int = 8, on_gpu: bool = False, ctx=None):
    """affine.rs:400-517 -> AlignResultC or None.  on_gpu: the affine model's hypotheses on the
    device (ab_ransac_affine_gpu, what the aligner runs); same result."""
    m = np.ascontiguousarray(matches, np.float64).reshape(-1, 4)
    out = AlignResultC()
    found = C.c_int(0)
    args = (m.ctypes.data if m.size else None, m.shape[0], 0 if method == "affine" else 1, n_threads,
            C.addressof(out), C.byref(found))
    if on_gpu:
        ctx = ctx or default_context()
        ctx.check(lib().ab_ransac_affine_gpu(ctx.handle, *args))
    else:
        st = lib().ab_ransac_affine(*args)
        if st:
            raise AstroBurstError(st, "ab_ransac_affine")
    return out if found.value else None


# ---------------------------------------------------------------- phase correlation
@dataclass
class PhaseCorrelationResult:
    """core/alignment/phase_correlation.rs:15-20."""
    dx: float
    dy: float
    confidence: float


def phase_correlate(reference: np.ndarray, target: np.ndarray, ctx=None) -> PhaseCorrelationResult:
    """phase_correlation.rs:22-99 (both images are cropped to the common minimum size)."""
    from ._lib import PhaseResultC
    ctx = ctx or default_context()
    r = np.ascontiguousarray(reference, dtype=np.float32)
    t = np.ascontiguousarray(target, dtype=np.float32)
    res = PhaseResultC()
    ctx.check(lib().ab_phase_correlate(ctx.handle, r.ctypes.data, r.shape[0], r.shape[1],
                                       t.ctypes.data, t.shape[0], t.shape[1], C.byref(res)))
    return PhaseCorrelationResult(res.dx, res.dy, res.confidence)


def is_low_confidence(confidence: float) -> bool:
    """phase_correlation.rs:163-165."""
    return confidence < 2.0


def area_downsample(img: np.ndarray, out_rows: int, out_cols: int, ctx=None) -> np.ndarray:
    """core/alignment/downsample.rs:6-46."""
    ctx = ctx or default_context()
    a = np.ascontiguousarray(img, dtype=np.float32)
    out = np.empty((out_rows, out_cols), np.float32)
    ctx.check(lib().ab_area_downsample(ctx.handle, a.ctypes.data, a.shape[0], a.shape[1],
                                       out_rows, out_cols, out.ctypes.data))
    return out


# ---- core::analysis::subframe (subframe.rs:52-158)
SubframeWeightConfig = SubframeConfigC


def analyze_subframe(image, config: SubframeConfigC | None = None, ctx: Context | None = None) -> SubframeMetricsC:
    """analyze_subframe, subframe.rs:52-118 (the path strings stay with the caller)."""
    ctx = ctx or default_context()
    d = _f32(image)
    cfg = config or SubframeConfigC()
    out = SubframeMetricsC()
    ctx.check(lib().ab_analyze_subframe(ctx.handle, d.ctypes.data, d.shape[0], d.shape[1], C.addressof(cfg),
                                        C.addressof(out)))
    return out


def normalize_weights(metrics):
    """normalize_weights, subframe.rs:147-158, in place on a list of SubframeMetricsC."""
    arr = (SubframeMetricsC * len(metrics))(*metrics)
    lib().ab_normalize_weights(arr, len(metrics))
    for m, a in zip(metrics, arr):
        m.weight = a.weight
    return metrics
