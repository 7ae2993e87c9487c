"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.h): imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package `astroburst_b200` never imports this module.

Function names follow the reference's Rust names (src-tauri/src/core/...).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (g++ -O3 -ffp-contract=off -fopenmp)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class ImageStats(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("min", "max", "median", "mad", "sigma", "mean")] + [
        ("valid_count", C.c_uint64)
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class StfParams(C.Structure):
    _fields_ = [("shadow", C.c_double), ("midtone", C.c_double), ("highlight", C.c_double)]


class Star(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("x", "y", "flux", "fwhm", "eccentricity", "peak")] + [
        ("npix", C.c_uint64),
        ("snr", C.c_double),
    ]


class Affine(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("a", "b", "tx", "c", "d", "ty")]

    def as_tuple(self):
        return (self.a, self.b, self.tx, self.c, self.d, self.ty)


class AlignResult(C.Structure):
    _fields_ = [
        ("transform", Affine),
        ("matched_stars", C.c_uint64),
        ("inliers", C.c_uint64),
        ("residual_px", C.c_double),
        ("method", C.c_int32),
    ]


_lib = None
_f32p = C.POINTER(C.c_float)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_sigma_clip_combine.restype = C.c_float
        _lib.orc_exact_median_mut.restype = C.c_double
        _lib.orc_median_f32_mut.restype = C.c_float
        _lib.orc_exact_mad_mut.restype = C.c_float
        _lib.orc_catmull_rom.restype = C.c_double
        _lib.orc_catmull_rom.argtypes = [C.c_double]
        _lib.orc_bicubic_sample.restype = C.c_float
        _lib.orc_bilinear_sample.restype = C.c_float
        _lib.orc_detect_stars.restype = C.c_size_t
        _lib.orc_match_stars.restype = C.c_size_t
        _lib.orc_star_mask.restype = C.c_size_t
        _lib.orc_masked_median.restype = C.c_double
        _lib.orc_masked_stretch_with_mask.restype = C.c_size_t
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


# ---------------------------------------------------------------- stacking
def sigma_clip_combine(values, sigma_low=3.0, sigma_high=3.0, max_iter=5):
    v = _f32(values).copy()
    rej = C.c_uint32(0)
    m = lib().orc_sigma_clip_combine(
        _p(v), C.c_size_t(v.size), C.c_float(sigma_low), C.c_float(sigma_high),
        C.c_size_t(max_iter), C.byref(rej))
    return np.float32(m), int(rej.value)


def stack_images(frames, sigma_low=3.0, sigma_high=3.0, max_iterations=5):
    """frames: sequence of equal-shape 2-D f32 arrays (align=false). -> (image, rejected_pixels)"""
    fr = [_f32(f) for f in frames]
    if not fr:
        raise ValueError("No images to stack")
    rows, cols = fr[0].shape
    ptrs = (C.c_void_p * len(fr))(*[f.ctypes.data for f in fr])
    out = np.empty((rows, cols), np.float32)
    rej = C.c_uint64(0)
    rc = lib().orc_stack_images(
        ptrs, C.c_size_t(len(fr)), C.c_size_t(rows), C.c_size_t(cols), C.c_float(sigma_low),
        C.c_float(sigma_high), C.c_size_t(max_iterations), _p(out), C.byref(rej))
    if rc != 0:
        raise ValueError("No images to stack")
    return out, int(rej.value)


# ---------------------------------------------------------------- phase correlation
def area_downsample(img, out_rows, out_cols):
    a = _f32(img)
    out = np.empty((out_rows, out_cols), np.float32)
    lib().orc_area_downsample(_p(a), C.c_size_t(a.shape[0]), C.c_size_t(a.shape[1]),
                              C.c_size_t(out_rows), C.c_size_t(out_cols), _p(out))
    return out


def phase_correlate(reference, target):
    """-> (dx, dy, confidence), phase_correlation.rs:22-99."""
    r, t = _f32(reference), _f32(target)
    dx, dy, conf = C.c_double(0), C.c_double(0), C.c_double(0)
    lib().orc_phase_correlate(_p(r), C.c_size_t(r.shape[0]), C.c_size_t(r.shape[1]), _p(t),
                              C.c_size_t(t.shape[0]), C.c_size_t(t.shape[1]), C.byref(dx),
                              C.byref(dy), C.byref(conf))
    return dx.value, dy.value, conf.value


def stack_images_aligned(frames, sigma_low=3.0, sigma_high=3.0, max_iterations=5):
    """stack_images with config.align = true (combine.rs:118-147): -> (image, rejected, offsets)."""
    fr = [_f32(f) for f in frames]
    rows = min(f.shape[0] for f in fr)
    cols = min(f.shape[1] for f in fr)
    fr = [np.ascontiguousarray(f[:rows, :cols]) for f in fr]
    aligned, offsets = [fr[0]], [(0, 0)]
    for f in fr[1:]:
        dx, dy, _ = phase_correlate(fr[0], f)
        aligned.append(shift_image_subpixel(f, dy, dx))
        rnd = lambda v: int(np.sign(v) * np.floor(abs(v) + 0.5))  # f64::round: half away from zero
        offsets.append((rnd(dy), rnd(dx)))
    img, rej = stack_images(aligned, sigma_low, sigma_high, max_iterations)
    return img, rej, offsets


# ---------------------------------------------------------------- FITS data-unit codec
_BPP = {8: 1, 16: 2, 32: 4, -32: 4, -64: 8}


def fits_decode(data: bytes, bitpix: int, bscale=1.0, bzero=0.0):
    raw = np.frombuffer(data, dtype=np.uint8)
    n = raw.size // _BPP[bitpix]
    out = np.empty(n, np.float32)
    lib().orc_fits_decode(_p(raw), C.c_size_t(n), C.c_int(bitpix), C.c_double(bscale), C.c_double(bzero), _p(out))
    return out


def fits_encode(data, bitpix: int, bzero=0.0, bscale=1.0) -> bytes:
    a = _f32(data).ravel()
    out = np.empty(a.size * _BPP[bitpix], np.uint8)
    lib().orc_fits_encode(_p(a), C.c_size_t(a.size), C.c_int(bitpix), C.c_double(bzero), C.c_double(bscale), _p(out))
    return out.tobytes()


# ---------------------------------------------------------------- calibration
def median_combine(frames, bias=None, dark=None, dark_ratio=1.0):
    """median_combine_row_major with the masters' per-frame preprocessing (calibration.rs:85-229)."""
    fr = [_f32(f) for f in frames]
    if not fr:
        raise ValueError("No frames provided")
    rows, cols = fr[0].shape
    ptrs = (C.c_void_p * len(fr))(*[f.ctypes.data for f in fr])
    b = _f32(bias) if bias is not None else None
    d = _f32(dark) if dark is not None else None
    out = np.empty((rows, cols), np.float32)
    lib().orc_median_combine(ptrs, C.c_size_t(len(fr)), C.c_size_t(rows), C.c_size_t(cols),
                             _p(b) if b is not None else None, _p(d) if d is not None else None,
                             C.c_float(dark_ratio), _p(out))
    return out


def sigma_clipped_mean_stack(frames, sigma_low=2.5, sigma_high=3.0, max_iterations=5):
    """calibration_pipeline.rs:317-378 -> (image, per-frame rejection counts)."""
    fr = [_f32(f) for f in frames]
    if not fr:
        raise ValueError("no frames")
    rows, cols = fr[0].shape
    ptrs = (C.c_void_p * len(fr))(*[f.ctypes.data for f in fr])
    out = np.empty((rows, cols), np.float32)
    counts = np.zeros(len(fr), np.uint64)
    lib().orc_sigma_clipped_mean_stack(ptrs, C.c_size_t(len(fr)), C.c_size_t(rows), C.c_size_t(cols),
                                       C.c_float(sigma_low), C.c_float(sigma_high),
                                       C.c_size_t(max_iterations), _p(out), _p(counts))
    return out, [int(c) for c in counts]


def normalize_flat(data):
    out = _f32(data).copy()
    lib().orc_normalize_flat(_p(out), C.c_size_t(out.size))
    return out


def calibrate_image(raw, bias=None, dark=None, flat=None, dark_ratio=1.0, clamp_negative=True):
    r = _f32(raw)
    b = _f32(bias) if bias is not None else None
    d = _f32(dark) if dark is not None else None
    f = _f32(flat) if flat is not None else None
    out = np.empty_like(r)
    lib().orc_calibrate_image(_p(r), C.c_size_t(r.size), _p(b) if b is not None else None,
                              _p(d) if d is not None else None, _p(f) if f is not None else None,
                              C.c_float(dark_ratio), C.c_int(1 if clamp_negative else 0), _p(out))
    return out


def exact_median_mut(data):
    v = _f32(data).copy()
    return lib().orc_exact_median_mut(_p(v), C.c_size_t(v.size))


def median_f32_mut(data):
    v = _f32(data).copy()
    return np.float32(lib().orc_median_f32_mut(_p(v), C.c_size_t(v.size)))


def exact_mad_mut(data, median):
    v = _f32(data).copy()
    return np.float32(lib().orc_exact_mad_mut(_p(v), C.c_size_t(v.size), C.c_float(median)))


def sigma_clipped_stats(values, kappa, iterations):
    v = _f32(values).copy()
    n = C.c_size_t(v.size)
    med, sig = C.c_double(), C.c_double()
    lib().orc_sigma_clipped_stats(_p(v), C.byref(n), C.c_float(kappa), C.c_size_t(iterations),
                                  C.byref(med), C.byref(sig))
    return med.value, sig.value


# ---------------------------------------------------------------- stats
def _stats_call(fn, data, *extra):
    d = _f32(data).reshape(-1)
    st = ImageStats()
    fn(_p(d), C.c_size_t(d.size), *extra, C.byref(st))
    return st


def compute_image_stats(data):
    return _stats_call(lib().orc_compute_image_stats, data)


def compute_image_stats_exact(data):
    return _stats_call(lib().orc_compute_image_stats_exact, data)


def compute_image_stats_hist(data):
    return _stats_call(lib().orc_compute_image_stats_hist, data)


def compute_image_stats_with_known_range(data, known_min, known_max):
    return _stats_call(lib().orc_compute_image_stats_with_known_range, data,
                       C.c_double(known_min), C.c_double(known_max))


def build_histogram(data, n_bins, dmin, dmax):
    d = _f32(data).reshape(-1)
    bins = np.zeros(n_bins, np.uint32)
    edges = np.zeros(n_bins + 1, np.float64)
    lib().orc_build_histogram(_p(d), C.c_size_t(d.size), C.c_size_t(n_bins), C.c_double(dmin),
                              C.c_double(dmax), _p(bins), _p(edges))
    return bins, edges


def downsample_histogram(bins, target_bins):
    b = np.ascontiguousarray(bins, np.uint32)
    out = np.zeros(min(target_bins, b.size), np.uint32)
    lib().orc_downsample_histogram(_p(b), C.c_size_t(b.size), C.c_size_t(target_bins), _p(out))
    return out


def auto_stf(stats: ImageStats, target_bg=0.25, shadow_k=-2.8) -> StfParams:
    out = StfParams()
    lib().orc_auto_stf(C.byref(stats), C.c_double(target_bg), C.c_double(shadow_k), C.byref(out))
    return out


def linked_stats(r: ImageStats, g: ImageStats, b: ImageStats) -> ImageStats:
    out = ImageStats()
    lib().orc_linked_stats(C.byref(r), C.byref(g), C.byref(b), C.byref(out))
    return out


def select_wb_reference(r: ImageStats, g: ImageStats, b: ImageStats):
    out = (C.c_double * 3)()
    lib().orc_select_wb_reference(C.byref(r), C.byref(g), C.byref(b), out)
    return tuple(out)


# ---------------------------------------------------------------- pointwise
def apply_stf(data, params: StfParams, stats: ImageStats):
    d = _f32(data)
    out = np.empty(d.shape, np.uint8)
    lib().orc_apply_stf_u8(_p(d), C.c_size_t(d.size), C.byref(params), C.byref(stats), _p(out))
    return out


def apply_stf_f32(data, params: StfParams, stats: ImageStats, rgb_variant=False):
    d = _f32(data)
    out = np.empty(d.shape, np.float32)
    fn = lib().orc_apply_stf_rgbvariant_f32 if rgb_variant else lib().orc_apply_stf_f32
    fn(_p(d), C.c_size_t(d.size), C.byref(params), C.byref(stats), _p(out))
    return out


def arcsinh_stretch_with_stats(data, dmin, dmax, factor, gamma=1.0):
    d = _f32(data)
    out = np.empty(d.shape, np.float32)
    lib().orc_arcsinh_stretch(_p(d), C.c_size_t(d.size), C.c_float(dmin), C.c_float(dmax),
                              C.c_float(factor), C.c_float(gamma), _p(out))
    return out


def blend_channels(channels, weights, rows, cols):
    """weights: list of (channel_idx, r, g, b)."""
    ch = [_f32(c).reshape(-1) for c in channels]
    ptrs = (C.c_void_p * max(len(ch), 1))(*[c.ctypes.data for c in ch])
    lens = (C.c_size_t * max(len(ch), 1))(*[c.size for c in ch])
    widx = np.array([w[0] for w in weights], np.uint64)
    wrgb = np.array([[w[1], w[2], w[3]] for w in weights], np.float64).reshape(-1)
    r = np.empty((rows, cols), np.float32)
    g = np.empty((rows, cols), np.float32)
    b = np.empty((rows, cols), np.float32)
    lib().orc_blend_channels(ptrs, lens, C.c_size_t(len(ch)), _p(widx), _p(wrgb),
                             C.c_size_t(len(weights)), C.c_size_t(rows), C.c_size_t(cols),
                             _p(r), _p(g), _p(b))
    return r, g, b


def apply_scnr_inplace(r, g, b, method="average", amount=1.0, preserve_luminance=False):
    """r,g,b: contiguous f32 arrays, modified in place (silent no-op on shape mismatch)."""
    if r.shape != g.shape or g.shape != b.shape:
        return
    for a in (r, g, b):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    lib().orc_apply_scnr(_p(r), _p(g), _p(b), C.c_size_t(r.size),
                         C.c_int(0 if method == "average" else 1), C.c_float(amount),
                         C.c_int(1 if preserve_luminance else 0))


def scale(data, factor):
    d = _f32(data)
    out = np.empty(d.shape, np.float32)
    lib().orc_scale(_p(d), C.c_size_t(d.size), C.c_float(factor), _p(out))
    return out


def apply_levels(data, black=0.0, gamma=1.0, white=1.0):
    d = _f32(data)
    out = np.empty(d.shape, np.float32)
    lib().orc_apply_levels(_p(d), C.c_size_t(d.size), C.c_double(black), C.c_double(gamma),
                           C.c_double(white), _p(out))
    return out


def spline_lut(points):
    pts = np.ascontiguousarray(points, np.float64).reshape(-1)
    lut = np.empty(4096, np.float32)
    lib().orc_spline_lut(_p(pts), C.c_size_t(pts.size // 2), _p(lut))
    return lut


def apply_curve(data, lut):
    d = _f32(data)
    l = _f32(lut)
    out = np.empty(d.shape, np.float32)
    lib().orc_apply_curve(_p(d), C.c_size_t(d.size), _p(l), _p(out))
    return out


# ---------------------------------------------------------------- resample
def catmull_rom(t):
    return lib().orc_catmull_rom(float(t))


def bicubic_sample(img, y, x):
    d = _f32(img)
    return np.float32(lib().orc_bicubic_sample(_p(d), C.c_size_t(d.shape[0]),
                                               C.c_size_t(d.shape[1]), C.c_double(y),
                                               C.c_double(x)))


def bilinear_sample(img, y, x):
    d = _f32(img)
    return np.float32(lib().orc_bilinear_sample(_p(d), C.c_size_t(d.shape[0]),
                                                C.c_size_t(d.shape[1]), C.c_double(y),
                                                C.c_double(x)))


def warp_image(img, transform, out_rows, out_cols):
    d = _f32(img)
    t = transform if isinstance(transform, Affine) else Affine(*transform)
    out = np.empty((out_rows, out_cols), np.float32)
    lib().orc_warp_image(_p(d), C.c_size_t(d.shape[0]), C.c_size_t(d.shape[1]), C.byref(t),
                         C.c_size_t(out_rows), C.c_size_t(out_cols), _p(out))
    return out


def shift_image_subpixel(img, dy, dx):
    d = _f32(img)
    out = np.empty(d.shape, np.float32)
    lib().orc_shift_image_subpixel(_p(d), C.c_size_t(d.shape[0]), C.c_size_t(d.shape[1]),
                                   C.c_double(dy), C.c_double(dx), _p(out))
    return out


def resample_image(img, target_rows, target_cols):
    d = _f32(img)
    if target_rows == 0 or target_cols == 0:
        raise ValueError("Target dimensions must be > 0")
    out = np.empty((target_rows, target_cols), np.float32)
    lib().orc_resample_image(_p(d), C.c_size_t(d.shape[0]), C.c_size_t(d.shape[1]),
                             C.c_size_t(target_rows), C.c_size_t(target_cols), _p(out))
    return out


# ---------------------------------------------------------------- stars / align
def estimate_background(img, tile_size):
    d = _f32(img)
    med, sig = C.c_double(), C.c_double()
    lib().orc_estimate_background(_p(d), C.c_size_t(d.shape[0]), C.c_size_t(d.shape[1]),
                                  C.c_size_t(tile_size), C.byref(med), C.byref(sig))
    return med.value, sig.value


@dataclass
class DetectionResult:
    stars: list
    background_median: float
    background_sigma: float
    threshold_sigma: float
    image_width: int
    image_height: int


def detect_stars(img, sigma_threshold, max_out=1 << 20) -> DetectionResult:
    d = _f32(img)
    buf = (Star * max_out)()
    med, sig = C.c_double(), C.c_double()
    n = lib().orc_detect_stars(_p(d), C.c_size_t(d.shape[0]), C.c_size_t(d.shape[1]),
                               C.c_double(sigma_threshold), buf, C.c_size_t(max_out),
                               C.byref(med), C.byref(sig))
    stars = [
        dict(x=s.x, y=s.y, flux=s.flux, fwhm=s.fwhm, eccentricity=s.eccentricity, peak=s.peak,
             npix=int(s.npix), snr=s.snr)
        for s in buf[: min(n, max_out)]
    ]
    return DetectionResult(stars, med.value, sig.value, sigma_threshold, d.shape[1], d.shape[0])


class SubframeConfig(C.Structure):
    """SubframeWeightConfig defaults, subframe.rs:37-50."""
    _fields_ = [(n, C.c_double) for n in ("fwhm_weight", "eccentricity_weight", "snr_weight", "noise_weight",
                                          "max_fwhm", "max_eccentricity", "min_snr")] + [("min_stars", C.c_uint64)]

    def __init__(self, fwhm_weight=1.0, eccentricity_weight=0.5, snr_weight=1.0, noise_weight=0.3,
                 max_fwhm=8.0, max_eccentricity=0.7, min_snr=5.0, min_stars=5):
        super().__init__(fwhm_weight, eccentricity_weight, snr_weight, noise_weight, max_fwhm,
                         max_eccentricity, min_snr, min_stars)


class SubframeMetrics(C.Structure):
    _fields_ = [("star_count", C.c_uint64)] + [(n, C.c_double) for n in (
        "median_fwhm", "median_eccentricity", "median_snr", "background_median", "background_sigma",
        "noise_ratio", "weight")] + [("accepted", C.c_int32)]

    def as_tuple(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)


def analyze_subframe(img, config: SubframeConfig | None = None) -> SubframeMetrics:
    d = _f32(img)
    cfg = config or SubframeConfig()
    out = SubframeMetrics()
    lib().orc_analyze_subframe(_p(d), C.c_size_t(d.shape[0]), C.c_size_t(d.shape[1]), C.byref(cfg), C.byref(out))
    return out


def subframe_weight(fwhm, ecc, snr, noise, config: SubframeConfig | None = None) -> float:
    cfg = config or SubframeConfig()
    lib().orc_subframe_weight.restype = C.c_double
    return lib().orc_subframe_weight(C.c_double(fwhm), C.c_double(ecc), C.c_double(snr), C.c_double(noise),
                                     C.byref(cfg))


def normalize_weights(weights):
    m = (SubframeMetrics * len(weights))()
    for i, w in enumerate(weights):
        m[i].weight = w
    lib().orc_normalize_weights(m, C.c_size_t(len(weights)))
    return [x.weight for x in m]


def normalize_for_detection(img):
    d = _f32(img)
    out = np.empty(d.shape, np.float32)
    lib().orc_normalize_for_detection(_p(d), C.c_size_t(d.size), _p(out))
    return out


def align_channel_affine(ref, tgt, n_threads=8) -> AlignResult:
    r, t = _f32(ref), _f32(tgt)
    out = AlignResult()
    lib().orc_align_channel_affine(_p(r), _p(t), C.c_size_t(r.shape[0]), C.c_size_t(r.shape[1]),
                                   C.c_int(n_threads), C.byref(out))
    return out


def match_stars(ref_xy, tgt_xy, max_matches=256):
    r = np.ascontiguousarray(ref_xy, np.float64).reshape(-1, 2)
    t = np.ascontiguousarray(tgt_xy, np.float64).reshape(-1, 2)
    out = np.zeros((max_matches, 4), np.float64)
    n = lib().orc_match_stars(_p(r), C.c_size_t(r.shape[0]), _p(t), C.c_size_t(t.shape[0]),
                              _p(out), C.c_size_t(max_matches))
    return out[: min(n, max_matches)].copy()


def ransac(matches, method="affine", n_threads=8):
    m = np.ascontiguousarray(matches, np.float64).reshape(-1, 4)
    out = AlignResult()
    ok = lib().orc_ransac(_p(m), C.c_size_t(m.shape[0]), C.c_int(0 if method == "affine" else 1),
                          C.c_int(n_threads), C.byref(out))
    return out if ok else None


def fit_affine(matches):
    m = np.ascontiguousarray(matches, np.float64).reshape(-1, 4)
    out = Affine()
    ok = lib().orc_fit_affine(_p(m), C.c_size_t(m.shape[0]), C.byref(out))
    return out if ok else None


def fit_rigid(matches):
    m = np.ascontiguousarray(matches, np.float64).reshape(-1, 4)
    out = Affine()
    ok = lib().orc_fit_rigid(_p(m), C.c_size_t(m.shape[0]), C.byref(out))
    return out if ok else None


# ---------------------------------------------------------------- preview tail / LRGB / masks (N4)
def preview_size(rows, cols, max_dim):
    pw, ph = C.c_size_t(), C.c_size_t()
    lib().orc_preview_size(C.c_size_t(rows), C.c_size_t(cols), C.c_size_t(max_dim), C.byref(pw),
                           C.byref(ph))
    return pw.value, ph.value


def render_rgb_u8(r, g, b, max_dim=0):
    """render_rgb / render_rgb_preview -> uint8 [ph, pw, 3]."""
    r, g, b = _f32(r), _f32(g), _f32(b)
    pw, ph = preview_size(r.shape[0], r.shape[1], max_dim)
    out = np.empty((ph, pw, 3), np.uint8)
    lib().orc_render_rgb_u8(_p(r), _p(g), _p(b), C.c_size_t(r.shape[0]), C.c_size_t(r.shape[1]),
                            C.c_size_t(max_dim), _p(out))
    return out


def render_rgb_u16be(r, g, b):
    """render_rgb_16bit's byte stream -> uint8 [rows, cols, 6]."""
    r, g, b = _f32(r), _f32(g), _f32(b)
    out = np.empty(r.shape + (6,), np.uint8)
    lib().orc_render_rgb_u16be(_p(r), _p(g), _p(b), C.c_size_t(r.size), _p(out))
    return out


def render_rgb_stf_u8(r, g, b, stf3, stats3, max_dim=0):
    r, g, b = _f32(r), _f32(g), _f32(b)
    pw, ph = preview_size(r.shape[0], r.shape[1], max_dim)
    out = np.empty((ph, pw, 3), np.uint8)
    sp = (StfParams * 3)(*stf3)
    st = (ImageStats * 3)(*stats3)
    lib().orc_render_rgb_stf_u8(_p(r), _p(g), _p(b), C.c_size_t(r.shape[0]), C.c_size_t(r.shape[1]),
                                sp, st, C.c_size_t(max_dim), _p(out))
    return out


def apply_lrgb(l, r, g, b, lightness_weight=1.0, chrominance_weight=1.0):
    """lrgb.rs:4-46: in place on r, g, b (float32 C-contiguous); raises on a shape mismatch."""
    if not (l.shape == r.shape == g.shape == b.shape):
        raise ValueError("L dimensions do not match RGB")
    lib().orc_apply_lrgb(_p(_f32(l)), _p(r), _p(g), _p(b), C.c_size_t(r.size),
                         C.c_float(lightness_weight), C.c_float(chrominance_weight))


def synthesize_luminance(r, g, b, finite_only=False):
    r, g, b = _f32(r), _f32(g), _f32(b)
    out = np.empty(r.shape, np.float32)
    lib().orc_luminance(_p(r), _p(g), _p(b), C.c_size_t(r.size), C.c_int(1 if finite_only else 0),
                        _p(out))
    return out


def star_mask_from_stars(img, stars, growth_factor=2.5, softness=4.0, min_fwhm=1.5, max_fwhm=30.0,
                         luminance_protect=False, luminance_ceiling=0.85):
    """generate_star_mask_from_detection -> (mask, stars_masked, coverage_fraction).
    stars: list of dicts with x, y, fwhm (detect_stars().stars)."""
    d = _f32(img)
    buf = (Star * max(len(stars), 1))()
    for i, s in enumerate(stars):
        buf[i].x, buf[i].y, buf[i].fwhm = s["x"], s["y"], s["fwhm"]
    mask = np.empty(d.shape, np.float32)
    cov = C.c_double()
    n = lib().orc_star_mask(_p(d), C.c_size_t(d.shape[0]), C.c_size_t(d.shape[1]), buf,
                            C.c_size_t(len(stars)), C.c_double(growth_factor), C.c_double(softness),
                            C.c_double(min_fwhm), C.c_double(max_fwhm),
                            C.c_int(1 if luminance_protect else 0), C.c_double(luminance_ceiling),
                            _p(mask), C.byref(cov))
    return mask, int(n), cov.value


def generate_star_mask(img, growth_factor=2.5, softness=4.0, detection_sigma=5.0, min_fwhm=1.5,
                       max_fwhm=30.0, luminance_protect=False, luminance_ceiling=0.85):
    det = detect_stars(img, detection_sigma)
    return star_mask_from_stars(img, det.stars, growth_factor, softness, min_fwhm, max_fwhm,
                                luminance_protect, luminance_ceiling)


def masked_median(img, mask):
    d, m = _f32(img), _f32(mask)
    return lib().orc_masked_median(_p(d), _p(m), C.c_size_t(d.size))


def masked_stretch_with_mask(img, mask, iterations=10, target_background=0.25,
                             protection_amount=0.85, convergence_threshold=1e-5):
    """-> (image, iterations_run, final_background, converged)."""
    d, m = _f32(img), _f32(mask)
    out = np.empty(d.shape, np.float32)
    fb, conv = C.c_double(), C.c_int()
    n = lib().orc_masked_stretch_with_mask(_p(d), C.c_size_t(d.size), _p(m), C.c_size_t(iterations),
                                           C.c_double(target_background),
                                           C.c_double(protection_amount),
                                           C.c_double(convergence_threshold), _p(out),
                                           C.byref(fb), C.byref(conv))
    return out, int(n), fb.value, bool(conv.value)


def masked_stretch(img, iterations=10, target_background=0.25, mask_growth=2.5, mask_softness=4.0,
                   luminance_protect=True, luminance_ceiling=0.85, protection_amount=0.85,
                   convergence_threshold=1e-5):
    """masked_stretch.rs:45-57 -> (image, iterations_run, final_background, converged,
    stars_masked, mask_coverage)."""
    mask, n, cov = generate_star_mask(img, mask_growth, mask_softness, 5.0, 1.5, 30.0,
                                      luminance_protect, luminance_ceiling)
    out, it, fb, conv = masked_stretch_with_mask(img, mask, iterations, target_background,
                                                 protection_amount, convergence_threshold)
    return out, it, fb, conv, n, cov
