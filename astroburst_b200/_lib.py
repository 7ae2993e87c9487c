"""ctypes loader for libastroburst_cuda.so and the Context wrapper (no torch types cross the ABI)."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# AB_LIBRARY: developer override (A/B-testing a differently built copy of the same library)
_SO = os.environ.get("AB_LIBRARY") or os.path.join(_HERE, "lib", "libastroburst_cuda.so")

AB_OK, AB_ERR_INVALID, AB_ERR_CUDA, AB_ERR_UNSUPPORTED, AB_ERR_NOMEM = range(5)
AB_MAX_FRAMES = 256


class AstroBurstError(RuntimeError):
    """Raised for every non-zero ab_status (the reference's anyhow::Error -> String)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[ab_status {status}] {message}")
        self.status = status
        self.message = message


class StackConfigC(C.Structure):
    _fields_ = [("sigma_low", C.c_float), ("sigma_high", C.c_float),
                ("max_iterations", C.c_uint32), ("align", C.c_uint32)]


class PhaseResultC(C.Structure):
    _fields_ = [("dx", C.c_double), ("dy", C.c_double), ("confidence", C.c_double)]


class CalibrationConfigC(C.Structure):
    _fields_ = [("master_bias", C.c_void_p), ("master_dark", C.c_void_p),
                ("master_flat", C.c_void_p), ("dark_exposure_ratio", C.c_float),
                ("clamp_negative", C.c_uint32)]


class ImageStatsC(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("min", "max", "median", "mad", "sigma", "mean")] + [
        ("valid_count", C.c_uint64)]


class StfParamsC(C.Structure):
    _fields_ = [("shadow", C.c_double), ("midtone", C.c_double), ("highlight", C.c_double)]


class AutoStfConfigC(C.Structure):
    _fields_ = [("target_bg", C.c_double), ("shadow_k", C.c_double)]


class ScnrConfigC(C.Structure):
    _fields_ = [("method", C.c_int32), ("amount", C.c_float), ("preserve_luminance", C.c_int32)]


class BlendWeightC(C.Structure):
    _fields_ = [("channel_idx", C.c_uint64), ("r_weight", C.c_double), ("g_weight", C.c_double),
                ("b_weight", C.c_double)]


class LevelsParamsC(C.Structure):
    _fields_ = [("black", C.c_double), ("gamma", C.c_double), ("white", C.c_double)]


class AffineC(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("a", "b", "tx", "c", "d", "ty")]


class StarC(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("x", "y", "flux", "fwhm", "eccentricity", "peak")] + [
        ("npix", C.c_uint64), ("snr", C.c_double)]


class DetectionInfoC(C.Structure):
    _fields_ = [("background_median", C.c_double), ("background_sigma", C.c_double),
                ("threshold_sigma", C.c_double), ("image_width", C.c_uint64),
                ("image_height", C.c_uint64), ("n_stars", C.c_uint64)]


class SubframeConfigC(C.Structure):
    """SubframeWeightConfig, core/analysis/subframe.rs:24-50 (defaults in __init__)."""
    _fields_ = [(n, C.c_double) for n in ("fwhm_weight", "eccentricity_weight", "snr_weight",
                                          "noise_weight", "max_fwhm", "max_eccentricity", "min_snr")] + [
        ("min_stars", C.c_uint64)]

    def __init__(self, fwhm_weight=1.0, eccentricity_weight=0.5, snr_weight=1.0, noise_weight=0.3,
                 max_fwhm=8.0, max_eccentricity=0.7, min_snr=5.0, min_stars=5):
        super().__init__(fwhm_weight, eccentricity_weight, snr_weight, noise_weight, max_fwhm,
                         max_eccentricity, min_snr, min_stars)


class SubframeMetricsC(C.Structure):
    """SubframeMetrics minus the path strings, subframe.rs:9-22."""
    _fields_ = [("star_count", C.c_uint64)] + [(n, C.c_double) for n in (
        "median_fwhm", "median_eccentricity", "median_snr", "background_median", "background_sigma",
        "noise_ratio", "weight")] + [("accepted", C.c_int32)]

    def as_tuple(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)


class AlignResultC(C.Structure):
    _fields_ = [("transform", AffineC), ("matched_stars", C.c_uint64), ("inliers", C.c_uint64),
                ("residual_px", C.c_double), ("method", C.c_int32)]


class ToneChainC(C.Structure):
    _fields_ = [("stf", StfParamsC * 3), ("stats", ImageStatsC * 3), ("levels", LevelsParamsC * 3),
                ("lut", C.c_void_p * 3), ("apply_scnr", C.c_int32), ("scnr", ScnrConfigC)]


class StarMaskConfigC(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("growth_factor", "softness", "detection_sigma", "min_fwhm",
                                          "max_fwhm")] + [("luminance_protect", C.c_int32),
                                                          ("luminance_ceiling", C.c_double)]


class StarMaskInfoC(C.Structure):
    _fields_ = [("stars_masked", C.c_uint64), ("coverage_fraction", C.c_double)]


class MaskedStretchConfigC(C.Structure):
    _fields_ = [("iterations", C.c_uint64), ("target_background", C.c_double),
                ("mask_growth", C.c_double), ("mask_softness", C.c_double),
                ("luminance_protect", C.c_int32), ("luminance_ceiling", C.c_double),
                ("protection_amount", C.c_double), ("convergence_threshold", C.c_double)]


class MaskedStretchInfoC(C.Structure):
    _fields_ = [("iterations_run", C.c_uint64), ("final_background", C.c_double),
                ("stars_masked", C.c_uint64), ("mask_coverage", C.c_double),
                ("converged", C.c_int32)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int)
ALLREDUCE_DEV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p)


def library_path() -> str:
    return _SO


_lib = None
_lib_lock = threading.Lock()

# name -> (restype, argtypes); the test-suite checks this table against include/astroburst_cuda.h
_vp, _sz, _u64p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)
SIGNATURES = {
    "ab_abi_version": (C.c_int, []),
    "ab_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "ab_ctx_destroy": (None, [_vp]),
    "ab_last_error": (C.c_char_p, [_vp]),
    "ab_ctx_device": (C.c_int, [_vp]),
    "ab_ctx_launch_count": (C.c_uint64, [_vp]),
    "ab_ctx_synchronize": (C.c_int, [_vp]),
    "ab_ctx_enable_peer_access": (C.c_int, [_vp, C.c_int]),
    "ab_ipc_export": (C.c_int, [_vp, _vp, _vp]),
    "ab_ipc_open": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "ab_ipc_close": (C.c_int, [_vp, _vp]),
    "ab_peer_export": (C.c_int, [_vp, _vp]),
    "ab_peer_connect": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "ab_peer_disconnect": (C.c_int, [_vp]),
    "ab_host_alloc": (C.c_int, [_sz, C.POINTER(_vp)]),
    "ab_host_free": (C.c_int, [_vp]),
    "ab_dev_alloc": (C.c_int, [_vp, _sz, C.POINTER(_vp)]),
    "ab_dev_free": (C.c_int, [_vp, _vp]),
    "ab_memcpy_h2d": (C.c_int, [_vp, _vp, _vp, _sz]),
    "ab_memcpy_d2h": (C.c_int, [_vp, _vp, _vp, _sz]),
    "ab_stack_sigma_clip": (C.c_int, [_vp, C.POINTER(_vp), _sz, _sz, _sz,
                                      C.POINTER(StackConfigC), _vp, _u64p]),
    "ab_stack_fits_units": (C.c_int, [_vp, C.POINTER(_vp), _vp, _vp, _vp, _sz, _sz, _sz,
                                      C.POINTER(StackConfigC), _vp, _u64p]),
    "ab_stack_sigma_clip_dev": (C.c_int, [_vp, C.POINTER(_vp), _sz, _sz,
                                          C.POINTER(StackConfigC), _vp, _vp, _vp]),
    "ab_sigma_clip_combine": (C.c_int, [_vp, _vp, _sz, C.c_float, C.c_float, C.c_uint32,
                                        C.POINTER(C.c_float), C.POINTER(C.c_uint32)]),
    # frame-sharded stacks over peer memory
    "ab_row_band": (None, [_sz, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "ab_frame_owner": (C.c_int, [_sz, _sz, C.c_int]),
    "ab_stack_frame_sharded_dev": (C.c_int, [_vp, C.POINTER(_vp), C.c_int, C.c_int, _sz, _sz, _sz,
                                             C.POINTER(StackConfigC), _vp, C.POINTER(_vp), C.c_int, _vp]),
    "ab_reshard_frames_to_bands_dev": (C.c_int, [_vp, C.POINTER(_vp), C.c_int, C.c_int, _sz, _sz, _sz,
                                                 _vp, _vp]),
    # phase correlation / aligned stack
    "ab_phase_correlate": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _vp]),
    "ab_phase_correlate_dev": (C.c_int, [_vp, _vp, _vp, _sz, _sz, _vp, _vp]),
    "ab_area_downsample": (C.c_int, [_vp, _vp, _sz, _sz, _sz, _sz, _vp]),
    "ab_area_downsample_dev": (C.c_int, [_vp, _vp, _sz, _sz, _sz, _sz, _vp, _vp]),
    "ab_stack_images": (C.c_int, [_vp, C.POINTER(_vp), _sz, _sz, _sz, C.POINTER(StackConfigC), _vp,
                                  _u64p, _vp]),
    # FITS data unit
    "ab_fits_decode": (C.c_int, [_vp, _vp, _sz, C.c_int, C.c_double, C.c_double, _vp]),
    "ab_fits_decode_dev": (C.c_int, [_vp, _vp, _sz, C.c_int, C.c_double, C.c_double, _vp, _vp]),
    "ab_fits_encode": (C.c_int, [_vp, _vp, _sz, C.c_int, C.c_double, C.c_double, _vp]),
    "ab_fits_encode_dev": (C.c_int, [_vp, _vp, _sz, C.c_int, C.c_double, C.c_double, _vp, _vp]),
    # calibration
    "ab_median_combine": (C.c_int, [_vp, C.POINTER(_vp), _sz, _sz, _sz, _vp, _vp, C.c_float, _vp]),
    "ab_median_combine_dev": (C.c_int, [_vp, C.POINTER(_vp), _sz, _sz, _vp, _vp, C.c_float, _vp, _vp]),
    "ab_sigma_clipped_mean_stack": (C.c_int, [_vp, C.POINTER(_vp), _sz, _sz, _sz, C.c_float, C.c_float,
                                              C.c_uint32, _vp, _vp]),
    "ab_sigma_clipped_mean_stack_dev": (C.c_int, [_vp, C.POINTER(_vp), _sz, _sz, C.c_float, C.c_float,
                                                  C.c_uint32, _vp, _vp, _vp]),
    "ab_normalize_flat": (C.c_int, [_vp, _vp, _sz]),
    "ab_normalize_flat_dev": (C.c_int, [_vp, _vp, _sz, _vp]),
    "ab_calibrate_image": (C.c_int, [_vp, _vp, _sz, _vp, _vp]),
    "ab_calibrate_image_dev": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _vp]),
    # pointwise
    "ab_apply_stf_u8": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "ab_apply_stf_u8_dev": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "ab_apply_stf_f32": (C.c_int, [_vp, _vp, _sz, _vp, _vp, C.c_int, _vp]),
    "ab_apply_stf_f32_dev": (C.c_int, [_vp, _vp, _sz, _vp, _vp, C.c_int, _vp, _vp]),
    "ab_arcsinh_stretch": (C.c_int, [_vp, _vp, _sz, C.c_float, C.c_float, C.c_float, C.c_float, _vp]),
    "ab_arcsinh_stretch_dev": (C.c_int, [_vp, _vp, _sz, C.c_float, C.c_float, C.c_float, C.c_float,
                                         _vp, _vp]),
    "ab_scale": (C.c_int, [_vp, _vp, _sz, C.c_float, _vp]),
    "ab_scale_dev": (C.c_int, [_vp, _vp, _sz, C.c_float, _vp, _vp]),
    "ab_apply_levels": (C.c_int, [_vp, _vp, _sz, _vp, _vp]),
    "ab_apply_levels_dev": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "ab_spline_lut": (C.c_int, [_vp, _sz, _vp]),
    "ab_apply_curve": (C.c_int, [_vp, _vp, _sz, _vp, _vp]),
    "ab_apply_curve_dev": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "ab_blend_channels": (C.c_int, [_vp, _vp, _vp, _sz, _vp, _sz, _sz, _sz, _vp, _vp, _vp]),
    "ab_blend_channels_dev": (C.c_int, [_vp, _vp, _vp, _sz, _vp, _sz, _sz, _sz, _vp, _vp, _vp, _vp]),
    "ab_apply_scnr": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _vp]),
    "ab_apply_scnr_dev": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "ab_tone_composite": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "ab_tone_composite_dev": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    # stats
    "ab_compute_image_stats": (C.c_int, [_vp, _vp, _sz, C.c_int, C.c_double, C.c_double, _vp]),
    "ab_image_stats_dev": (C.c_int, [_vp, _vp, _sz, C.c_int, C.c_double, C.c_double, _vp, _vp]),
    "ab_image_stats_sharded_dev": (C.c_int, [_vp, _vp, _sz, _vp, _vp, C.c_int, C.c_double, C.c_double,
                                             _vp, _vp]),
    "ab_image_stats_sharded_chained_dev": (C.c_int, [_vp, _vp, _sz, _vp, _vp, C.c_int, C.c_double,
                                                     C.c_double, _vp, _vp]),
    "ab_image_stats_peer_dev": (C.c_int, [_vp, _vp, _sz, C.c_int, C.c_double, C.c_double, _vp, _vp]),
    "ab_histogram_u32": (C.c_int, [_vp, _vp, _sz, _sz, C.c_double, C.c_double, _vp, _vp]),
    "ab_histogram_u32_dev": (C.c_int, [_vp, _vp, _sz, _sz, C.c_double, C.c_double, _vp, _vp, _vp]),
    "ab_downsample_histogram": (C.c_int, [_vp, _sz, _sz, _vp]),
    "ab_auto_stf": (C.c_int, [_vp, _vp, _vp]),
    "ab_linked_stats": (C.c_int, [_vp, _vp, _vp, _vp]),
    "ab_select_wb_reference": (C.c_int, [_vp, _vp, _vp, _vp]),
    # bicubic gathers
    "ab_warp_affine_bicubic": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _vp]),
    "ab_warp_affine_bicubic_dev": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _vp, _vp]),
    "ab_shift_bicubic": (C.c_int, [_vp, _vp, _sz, _sz, C.c_double, C.c_double, _vp]),
    "ab_shift_bicubic_dev": (C.c_int, [_vp, _vp, _sz, _sz, C.c_double, C.c_double, _vp, _vp]),
    "ab_resample_bicubic": (C.c_int, [_vp, _vp, _sz, _sz, _sz, _sz, _vp]),
    "ab_resample_bicubic_dev": (C.c_int, [_vp, _vp, _sz, _sz, _sz, _sz, _vp, _vp]),
    # stars / alignment
    "ab_detect_stars": (C.c_int, [_vp, _vp, _sz, _sz, C.c_double, _vp, _sz, _vp]),
    "ab_detect_stars_dev": (C.c_int, [_vp, _vp, _sz, _sz, C.c_double, _vp, _sz, _vp, _vp]),
    "ab_analyze_subframe": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _vp]),
    "ab_analyze_subframe_dev": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _vp, _vp]),
    "ab_normalize_weights": (None, [_vp, _sz]),
    "ab_estimate_background": (C.c_int, [_vp, _vp, _sz, _sz, _sz, C.POINTER(C.c_double),
                                         C.POINTER(C.c_double)]),
    "ab_normalize_for_detection": (C.c_int, [_vp, _vp, _sz, _vp]),
    "ab_align_affine": (C.c_int, [_vp, _vp, _vp, _sz, _sz, C.c_int, _vp]),
    "ab_align_affine_dev": (C.c_int, [_vp, _vp, _vp, _sz, _sz, C.c_int, _vp, _vp]),
    "ab_align_ref_stars_dev": (C.c_int, [_vp, _vp, _sz, _sz, _vp, C.POINTER(C.c_size_t), _vp]),
    "ab_align_affine_with_ref_dev": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _sz, _sz, C.c_int, _vp, _vp]),
    "ab_align_affine_from_stars_dev": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _vp, _vp, _sz, _sz, C.c_int, _vp, _vp]),
    "ab_match_stars": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, C.POINTER(C.c_size_t)]),
    "ab_ransac_affine": (C.c_int, [_vp, _sz, C.c_int, C.c_int, _vp, C.POINTER(C.c_int)]),
    "ab_ransac_affine_gpu": (C.c_int, [_vp, _vp, _sz, C.c_int, C.c_int, _vp, C.POINTER(C.c_int)]),
    # preview / export tail, LRGB, star mask, masked stretch
    "ab_preview_size": (C.c_int, [_sz, _sz, _sz, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "ab_render_rgb_u8": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _sz, _sz, _vp]),
    "ab_render_rgb_u8_dev": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _sz, _sz, _vp, _vp]),
    "ab_render_rgb_stf_u8": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _sz, _vp, _vp, _sz, _vp]),
    "ab_render_rgb_stf_u8_dev": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _sz, _vp, _vp, _sz, _vp, _vp]),
    "ab_render_rgb_u16be": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _vp]),
    "ab_render_rgb_u16be_dev": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "ab_apply_lrgb": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _sz, C.c_float, C.c_float]),
    "ab_apply_lrgb_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _sz, C.c_float, C.c_float, _vp]),
    "ab_luminance": (C.c_int, [_vp, _vp, _vp, _vp, _sz, C.c_int, _vp]),
    "ab_luminance_dev": (C.c_int, [_vp, _vp, _vp, _vp, _sz, C.c_int, _vp, _vp]),
    "ab_star_mask_from_stars": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _sz, _vp, _vp, _vp]),
    "ab_star_mask_from_stars_dev": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _sz, _vp, _vp, _vp, _vp]),
    "ab_generate_star_mask": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _vp, _vp]),
    "ab_generate_star_mask_dev": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _vp, _vp, _vp]),
    "ab_masked_median_dev": (C.c_int, [_vp, _vp, _vp, _sz, C.POINTER(C.c_double), _vp]),
    "ab_masked_stretch_with_mask": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "ab_masked_stretch_with_mask_dev": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "ab_masked_stretch": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _vp, _vp]),
    "ab_masked_stretch_dev": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _vp, _vp, _vp, _vp]),
}


def lib() -> C.CDLL:
    """Load the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(_SO):
                raise AstroBurstError(
                    AB_ERR_CUDA,
                    f"{_SO} is missing: build it with `python -m astroburst_b200.build` "
                    "(there is no CPU fallback)")
            l = C.CDLL(_SO)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(l, name)
                fn.restype = res
                fn.argtypes = args
            _lib = l
    return _lib


class Context:
    """Owns one `ab_ctx` (one CUDA device, its streams and scratch)."""

    def __init__(self, device: int | None = None):
        self._h = C.c_void_p()
        st = lib().ab_ctx_create(-1 if device is None else int(device), C.byref(self._h))
        if st != AB_OK:
            self._h = C.c_void_p()
            raise AstroBurstError(st, "ab_ctx_create failed (no usable CUDA device?)")

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    @property
    def device(self) -> int:
        return lib().ab_ctx_device(self._h)

    @property
    def launch_count(self) -> int:
        return int(lib().ab_ctx_launch_count(self._h))

    def check(self, status: int) -> None:
        if status != AB_OK:
            msg = lib().ab_last_error(self._h)
            raise AstroBurstError(status, msg.decode() if msg else "")

    def synchronize(self) -> None:
        self.check(lib().ab_ctx_synchronize(self._h))

    def close(self) -> None:
        if self._h:
            lib().ab_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default = {}
_default_lock = threading.Lock()


def default_context(device: int | None = None) -> Context:
    """Process-wide context per device (created on first use)."""
    key = -1 if device is None else int(device)
    with _default_lock:
        if key not in _default:
            _default[key] = Context(device)
        return _default[key]
