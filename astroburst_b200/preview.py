"""Host-side mirror of the preview/export tail, the LRGB merge, the star mask and the masked
stretch over the C ABI.  Names and argument meaning follow the Rust functions
they replace (file:line cited per function, paths relative to src-tauri/src/)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import (ImageStatsC, MaskedStretchConfigC, MaskedStretchInfoC, StarC, StarMaskConfigC,
                   StarMaskInfoC, StfParamsC, default_context, lib)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def preview_size(rows: int, cols: int, max_dim: int):
    """cmd/helpers.rs:222-227 -> (pw, ph); max_dim = 0 keeps the full size."""
    pw, ph = C.c_size_t(), C.c_size_t()
    lib().ab_preview_size(rows, cols, max_dim, C.byref(pw), C.byref(ph))
    return pw.value, ph.value


def render_rgb(r, g, b, max_dim: int = 0, ctx=None) -> np.ndarray:
    """render_rgb (infra/render/rgb.rs:7-47) / render_rgb_preview (cmd/helpers.rs:204-262) without
    the PNG encoder -> uint8 [ph, pw, 3]."""
    ctx = ctx or default_context()
    r, g, b = _f32(r), _f32(g), _f32(b)
    if not (r.shape == g.shape == b.shape) or r.ndim != 2:
        raise ValueError("R, G, B must be 2-D planes of one shape")
    pw, ph = preview_size(r.shape[0], r.shape[1], max_dim)
    out = np.empty((ph, pw, 3), np.uint8)
    ctx.check(lib().ab_render_rgb_u8(ctx.handle, r.ctypes.data, g.ctypes.data, b.ctypes.data,
                                     r.shape[0], r.shape[1], max_dim, out.ctypes.data))
    return out


def render_rgb_16bit(r, g, b, ctx=None) -> np.ndarray:
    """render_rgb_16bit, rgb.rs:49-98: the big-endian byte stream handed to the PNG encoder
    -> uint8 [rows, cols, 6]."""
    ctx = ctx or default_context()
    r, g, b = _f32(r), _f32(g), _f32(b)
    out = np.empty(r.shape + (6,), np.uint8)
    ctx.check(lib().ab_render_rgb_u16be(ctx.handle, r.ctypes.data, g.ctypes.data, b.ctypes.data,
                                        r.size, out.ctypes.data))
    return out


def render_rgb_preview_with_stf(r, g, b, stf3, stats3, max_dim: int = 0, ctx=None) -> np.ndarray:
    """cmd/helpers.rs:264-322 with make_stf_u8_fn closures (stf.rs:122-145): stf3 / stats3 are the
    (r, g, b) StfParams / ImageStats."""
    ctx = ctx or default_context()
    r, g, b = _f32(r), _f32(g), _f32(b)
    pw, ph = preview_size(r.shape[0], r.shape[1], max_dim)
    out = np.empty((ph, pw, 3), np.uint8)
    sp = (StfParamsC * 3)(*stf3)
    st = (ImageStatsC * 3)(*stats3)
    ctx.check(lib().ab_render_rgb_stf_u8(ctx.handle, r.ctypes.data, g.ctypes.data, b.ctypes.data,
                                         r.shape[0], r.shape[1], sp, st, max_dim, out.ctypes.data))
    return out


def apply_lrgb(l, r, g, b, lightness_weight: float = 1.0, chrominance_weight: float = 1.0, ctx=None):
    """core/compose/lrgb.rs:4-46: in place on r, g, b (float32 C-contiguous arrays); raises on a
    dimension mismatch like the reference's bail! (:14-19)."""
    ctx = ctx or default_context()
    if not (l.shape == r.shape == g.shape == b.shape):
        raise ValueError(f"L dimensions {l.shape} do not match RGB (R: {r.shape}, G: {g.shape}, "
                         f"B: {b.shape})")
    for a in (r, g, b):
        if a.dtype != np.float32 or not a.flags.c_contiguous:
            raise ValueError("r, g, b must be float32 C-contiguous (updated in place)")
    l = _f32(l)
    ctx.check(lib().ab_apply_lrgb(ctx.handle, l.ctypes.data, r.ctypes.data, g.ctypes.data,
                                  b.ctypes.data, r.size, lightness_weight, chrominance_weight))


def synthesize_luminance(r, g, b, finite_only: bool = False, ctx=None) -> np.ndarray:
    """lrgb.rs:48-62; finite_only=True is masked_stretch.rs:134-159 compute_luminance."""
    ctx = ctx or default_context()
    r, g, b = _f32(r), _f32(g), _f32(b)
    if not (r.shape == g.shape == b.shape):
        raise ValueError(f"Channel dimension mismatch: R={r.shape} G={g.shape} B={b.shape}")
    out = np.empty(r.shape, np.float32)
    ctx.check(lib().ab_luminance(ctx.handle, r.ctypes.data, g.ctypes.data, b.ctypes.data, r.size,
                                 1 if finite_only else 0, out.ctypes.data))
    return out


@dataclass
class StarMaskConfig:  # star_mask.rs:7-29
    growth_factor: float = 2.5
    softness: float = 4.0
    detection_sigma: float = 5.0
    min_fwhm: float = 1.5
    max_fwhm: float = 30.0
    luminance_protect: bool = False
    luminance_ceiling: float = 0.85

    def c(self):
        return StarMaskConfigC(self.growth_factor, self.softness, self.detection_sigma, self.min_fwhm,
                               self.max_fwhm, 1 if self.luminance_protect else 0,
                               self.luminance_ceiling)


@dataclass
class StarMaskResult:  # star_mask.rs:31-36
    mask: np.ndarray
    stars_masked: int
    coverage_fraction: float


def generate_star_mask_from_detection(image, stars, config: StarMaskConfig | None = None,
                                      ctx=None) -> StarMaskResult:
    """star_mask.rs:46-138; stars: DetectionResult.stars (dicts or StarC with x, y, fwhm)."""
    ctx = ctx or default_context()
    config = config or StarMaskConfig()
    d = _f32(image)
    buf = (StarC * max(len(stars), 1))()
    for i, s in enumerate(stars):
        get = (lambda k: s[k]) if isinstance(s, dict) else (lambda k: getattr(s, k))
        buf[i].x, buf[i].y, buf[i].fwhm = get("x"), get("y"), get("fwhm")
    mask = np.empty(d.shape, np.float32)
    cfg, info = config.c(), StarMaskInfoC()
    ctx.check(lib().ab_star_mask_from_stars(ctx.handle, d.ctypes.data, d.shape[0], d.shape[1], buf,
                                            len(stars), C.addressof(cfg), mask.ctypes.data,
                                            C.addressof(info)))
    return StarMaskResult(mask, int(info.stars_masked), info.coverage_fraction)


def generate_star_mask(image, config: StarMaskConfig | None = None, ctx=None) -> StarMaskResult:
    """star_mask.rs:38-44."""
    ctx = ctx or default_context()
    config = config or StarMaskConfig()
    d = _f32(image)
    mask = np.empty(d.shape, np.float32)
    cfg, info = config.c(), StarMaskInfoC()
    ctx.check(lib().ab_generate_star_mask(ctx.handle, d.ctypes.data, d.shape[0], d.shape[1],
                                          C.addressof(cfg), mask.ctypes.data, C.addressof(info)))
    return StarMaskResult(mask, int(info.stars_masked), info.coverage_fraction)


@dataclass
class MaskedStretchConfig:  # masked_stretch.rs:7-32
    iterations: int = 10
    target_background: float = 0.25
    mask_growth: float = 2.5
    mask_softness: float = 4.0
    luminance_protect: bool = True
    luminance_ceiling: float = 0.85
    protection_amount: float = 0.85
    convergence_threshold: float = 1e-5

    def c(self):
        return MaskedStretchConfigC(self.iterations, self.target_background, self.mask_growth,
                                    self.mask_softness, 1 if self.luminance_protect else 0,
                                    self.luminance_ceiling, self.protection_amount,
                                    self.convergence_threshold)


@dataclass
class MaskedStretchResult:  # masked_stretch.rs:34-42
    image: np.ndarray
    iterations_run: int
    final_background: float
    stars_masked: int
    mask_coverage: float
    converged: bool


def _result(out, info, stars=None, cov=None):
    return MaskedStretchResult(out, int(info.iterations_run), info.final_background,
                               int(info.stars_masked) if stars is None else stars,
                               info.mask_coverage if cov is None else cov, bool(info.converged))


def masked_stretch_with_mask(image, mask_result: StarMaskResult,
                             config: MaskedStretchConfig | None = None, ctx=None) -> MaskedStretchResult:
    """masked_stretch.rs:59-117."""
    ctx = ctx or default_context()
    config = config or MaskedStretchConfig()
    d, m = _f32(image), _f32(mask_result.mask)
    if d.shape != m.shape:
        raise ValueError("mask and image dimensions differ")
    out = np.empty(d.shape, np.float32)
    cfg, info = config.c(), MaskedStretchInfoC()
    ctx.check(lib().ab_masked_stretch_with_mask(ctx.handle, d.ctypes.data, d.size, m.ctypes.data,
                                                C.addressof(cfg), out.ctypes.data, C.addressof(info)))
    return _result(out, info, mask_result.stars_masked, mask_result.coverage_fraction)


def masked_stretch(image, config: MaskedStretchConfig | None = None, ctx=None) -> MaskedStretchResult:
    """masked_stretch.rs:45-57."""
    ctx = ctx or default_context()
    config = config or MaskedStretchConfig()
    d = _f32(image)
    out = np.empty(d.shape, np.float32)
    cfg, info = config.c(), MaskedStretchInfoC()
    ctx.check(lib().ab_masked_stretch(ctx.handle, d.ctypes.data, d.shape[0], d.shape[1],
                                      C.addressof(cfg), out.ctypes.data, C.addressof(info)))
    return _result(out, info)


def masked_stretch_rgb_shared(r, g, b, config: MaskedStretchConfig | None = None, ctx=None):
    """masked_stretch.rs:161-190: one mask from the finite-only luminance, shared by the three
    channels -> (res_r, res_g, res_b, shared_mask_coverage, shared_stars_masked)."""
    config = config or MaskedStretchConfig()
    lum = synthesize_luminance(r, g, b, finite_only=True, ctx=ctx)
    mc = StarMaskConfig(growth_factor=config.mask_growth, softness=config.mask_softness,
                        luminance_protect=config.luminance_protect,
                        luminance_ceiling=config.luminance_ceiling)
    shared = generate_star_mask(lum, mc, ctx)
    res = [masked_stretch_with_mask(ch, shared, config, ctx) for ch in (r, g, b)]
    return res[0], res[1], res[2], shared.coverage_fraction, shared.stars_masked
