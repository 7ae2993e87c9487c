"""Device-resident compositions of the hot-path operators in the order the reference's commands fix
(SURVEY 8(a) A19), for BASELINE configs 3 and 4.  Inputs and outputs are device pointers / torch
tensors; every pixel sweep is one `_dev` entry point of the C ABI, every scalar step is the host
scalar code the reference runs between them.

config 4 (`color_chain`): blend_channels_cmd (cmd/compose/blend.rs:129-223) -> compute_auto_wb_cmd
(cmd/compose/color.rs:188-224) -> calibrate_and_scnr_cmd (color.rs:98-185) -> apply_tone_composite_cmd
with linked STF + curves (cmd/processing/curves.rs:58-191).  Works on a ROW BAND: the pointwise
sweeps need nothing from other bands, the statistics come from `stats_fn`, which is either the
single-GPU ab_image_stats_dev or the sharded, NCCL-reduced variant (parallel.py).

`process_fits`: the mono branch of process_fits / process_fits_full (cmd/io/mod.rs:23-31,103-170) from
the file's data unit to the u8 preview plane and the display histogram.

config 3 (`align_and_warp`): align_channels_cmd with alignMethod = "affine"
(cmd/compose/blend.rs:226-348): every target frame is registered against the reference frame
(align_channel_affine, affine.rs:129-212) and resampled with warp_image (affine.rs:663-690).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (AlignResultC, AutoStfConfigC, BlendWeightC, Context, ImageStatsC, LevelsParamsC, ScnrConfigC,
                   StfParamsC, ToneChainC, default_context, lib)

PAR_THRESHOLD = 4_000_000  # cmd/compose/color.rs:19


def stats_single_gpu(ctx: Context, stream: int):
    """stats_fn for an image that lives whole on one GPU: compute_image_stats /
    compute_image_stats_with_known_range (stats.rs:15-41)."""

    def fn(ptr: int, n: int, known_range=None) -> ImageStatsC:
        st = ImageStatsC()
        mode, (kmin, kmax) = (3, known_range) if known_range is not None else (0, (0.0, 0.0))
        ctx.check(lib().ab_image_stats_dev(ctx.handle, ptr, n, mode, kmin, kmax, C.addressof(st),
                                           stream if stream else None))
        return st

    return fn


def stats_sharded(ctx: Context, stream: int, allreduce_dev, total_pixels: int):
    """stats_fn for a row-band sharded image: histogram algorithm with the reductions
    on the device.  `total_pixels` decides, like the reference, between the exact and the histogram
    algorithm -- sharded images are far above the 4,000,000-pixel switch, anything else is refused."""
    if total_pixels <= PAR_THRESHOLD:
        raise ValueError("sharded statistics implement the histogram algorithm (> 4,000,000 px) only")

    def fn(ptr: int, n: int, known_range=None) -> ImageStatsC:
        from . import parallel

        return parallel.compute_image_stats_sharded_chained(ptr, n, allreduce_dev, known_range, stream, ctx)

    return fn


def stats_peer(ctx: Context, stream: int, total_pixels: int):
    """stats_fn for a row-band sharded image over the library's peer group (parallel.connect_peer_group):
    reductions inside kernels over NVLink, no NCCL, no host callback."""
    if total_pixels <= PAR_THRESHOLD:
        raise ValueError("sharded statistics implement the histogram algorithm (> 4,000,000 px) only")

    def fn(ptr: int, n: int, known_range=None) -> ImageStatsC:
        from . import parallel

        return parallel.compute_image_stats_peer(ptr, n, known_range, stream, ctx)

    return fn


def auto_wb_factors(sr: ImageStatsC, sg: ImageStatsC, sb: ImageStatsC):
    """compute_auto_wb_cmd, cmd/compose/color.rs:188-224 -> (r_factor, g_factor, b_factor)."""

    def stability(med, mad):
        return mad / med if med > 1e-10 else float(np.finfo(np.float64).max)

    stab_r, stab_g, stab_b = stability(sr.median, sr.mad), stability(sg.median, sg.mad), stability(sb.median, sb.mad)
    if stab_r <= stab_g and stab_r <= stab_b:
        m = max(sr.median, 1e-10)
        return 1.0, m / max(sg.median, 1e-10), m / max(sb.median, 1e-10)
    if stab_b <= stab_g:
        m = max(sb.median, 1e-10)
        return m / max(sr.median, 1e-10), m / max(sg.median, 1e-10), 1.0
    m = max(sg.median, 1e-10)
    return m / max(sr.median, 1e-10), 1.0, m / max(sb.median, 1e-10)


def _sync_stream(stream: int) -> None:
    import torch

    if stream:
        torch.cuda.ExternalStream(stream).synchronize()
    else:
        torch.cuda.synchronize()


@dataclass
class StatsLane:
    """One of the host threads that compute the three channel statistics of a colour-chain stage
    side by side: its own context (scratch, peer group), its stream and its stats_fn."""
    ctx: Context
    stream: int
    stats_fn: Callable


@dataclass
class ColorChainResult:
    stats_blend: tuple          # ImageStats of the blended R, G, B (blend.rs:187-193)
    factors: tuple              # white-balance factors as applied (f32, floored at 1e-6)
    stats_balanced: tuple       # after WB (+ SCNR re-stat of G, color.rs:142-155)
    stf: StfParamsC             # linked auto-STF (helpers.rs:185-202)
    launches: int
    lut: object = None          # device LUT of the curve: keep it alive until the stream has drained


def color_chain(ch_ptrs: Sequence[int], n: int, rows: int, cols: int, weights, curve_points,
                r_ptr: int, g_ptr: int, b_ptr: int, stats_fn: Callable, total_pixels: Optional[int] = None,
                scnr_after_wb=(0, 0.5, 0), ctx: Context | None = None, stream: int = 0,
                lanes: Optional[Sequence[StatsLane]] = None, pool=None) -> ColorChainResult:
    """The config-4 colour chain on `n` = rows * cols pixels (one row band, or the whole image).
    ch_ptrs: device planes of the input channels; r/g/b_ptr: three device planes that receive the
    result (they are also the working planes).  weights: [(channel_idx, r, g, b), ...].
    scnr_after_wb: (method, amount, preserve_luminance) of calibrate_and_scnr_cmd.
    lanes (three StatsLane) + pool (a 3-thread executor): the R, G, B statistics of a stage -- and
    the white-balance scale in front of them -- run side by side on three contexts; every statistic
    ends in a small read-back, and the other channels' sweeps fill the GPU meanwhile."""
    ctx = ctx or default_context()
    h = ctx.handle
    s = stream if stream else None
    all_ctx = list({id(c): c for c in [ctx] + ([l.ctx for l in lanes] if lanes else [])}.values())
    l0 = sum(c.launch_count for c in all_ctx)
    total_pixels = total_pixels or n
    use_lanes = bool(lanes) and pool is not None and len(lanes) == 3

    def per_channel(fn):
        """fn(k, lane) for k = 0, 1, 2 on the three lanes -> [result_r, result_g, result_b]"""
        futs = [pool.submit(fn, k, lanes[k]) for k in (1, 2)]
        first = fn(0, lanes[0])
        return [first] + [f.result() for f in futs]
    # ---- blend_channels_cmd: blend + 3 x compute_image_stats
    w = (BlendWeightC * len(weights))(*[BlendWeightC(*x) for x in weights])
    ptrs = (C.c_void_p * len(ch_ptrs))(*[int(p) for p in ch_ptrs])
    lens = (C.c_size_t * len(ch_ptrs))(*([n] * len(ch_ptrs)))
    ctx.check(lib().ab_blend_channels_dev(h, ptrs, lens, len(ch_ptrs), w, len(weights), rows, cols, r_ptr, g_ptr,
                                          b_ptr, s))
    planes = (r_ptr, g_ptr, b_ptr)
    if use_lanes:
        _sync_stream(stream)     # the blend has finished before the other lanes' streams read its planes
        s_r, s_g, s_b = per_channel(lambda k, lane: lane.stats_fn(planes[k], n))
    else:
        s_r, s_g, s_b = stats_fn(r_ptr, n), stats_fn(g_ptr, n), stats_fn(b_ptr, n)
    stats_blend = (s_r, s_g, s_b)
    # ---- compute_auto_wb_cmd -> calibrate_and_scnr_cmd
    fac = [max(np.float32(f), np.float32(1e-6)) for f in auto_wb_factors(s_r, s_g, s_b)]   # color.rs:115-117
    def calibrate_channel(ptr, f, st, c, cs, fn):   # calibrate_channel, color.rs:21-49
        c.check(lib().ab_scale_dev(c.handle, ptr, n, C.c_float(float(f)), ptr, cs if cs else None))
        if total_pixels <= PAR_THRESHOLD:
            return fn(ptr, n)
        # factor >= 0 here (floored at 1e-6): known range = orig range * factor (f64)
        return fn(ptr, n, (st.min * float(f), st.max * float(f)))

    before = (s_r, s_g, s_b)
    if use_lanes:
        bal = per_channel(lambda k, lane: calibrate_channel(planes[k], fac[k], before[k], lane.ctx, lane.stream,
                                                            lane.stats_fn))
    else:
        bal = [calibrate_channel(planes[k], fac[k], before[k], ctx, stream, stats_fn) for k in range(3)]
    method, amount, preserve = scnr_after_wb
    if amount > 1e-7:
        cfg = ScnrConfigC(method, amount, preserve)
        ctx.check(lib().ab_apply_scnr_dev(h, r_ptr, g_ptr, b_ptr, n, C.addressof(cfg), s))
        if preserve and use_lanes:
            _sync_stream(stream)
            bal = per_channel(lambda k, lane: lane.stats_fn(planes[k], n))
        elif preserve:
            bal = [stats_fn(r_ptr, n), stats_fn(g_ptr, n), stats_fn(b_ptr, n)]
        else:
            bal[1] = stats_fn(g_ptr, n)
    # ---- apply_tone_composite_cmd, linked STF: combined stats normalise all three channels
    combined = ImageStatsC()
    lib().ab_linked_stats(C.addressof(bal[0]), C.addressof(bal[1]), C.addressof(bal[2]), C.addressof(combined))
    cfg_stf = AutoStfConfigC(0.25, -2.8)
    stf = StfParamsC()
    lib().ab_auto_stf(C.addressof(combined), C.addressof(cfg_stf), C.addressof(stf))
    import torch

    lut_host = np.empty(4096, np.float32)
    pts = np.ascontiguousarray(curve_points, dtype=np.float64)
    ctx.check(lib().ab_spline_lut(pts.ctypes.data, len(pts), lut_host.ctypes.data))
    lut = torch.from_numpy(lut_host).to(torch.device("cuda", ctx.device))
    tc = ToneChainC()
    for k in range(3):
        tc.stf[k], tc.stats[k] = stf, combined
        tc.levels[k] = LevelsParamsC(0.0, 1.0, 1.0)
        tc.lut[k] = lut.data_ptr()
    tc.apply_scnr, tc.scnr = 0, ScnrConfigC(0, 0.0, 0)
    ctx.check(lib().ab_tone_composite_dev(h, r_ptr, g_ptr, b_ptr, n, C.addressof(tc), r_ptr, g_ptr, b_ptr, s))
    l1 = sum(c.launch_count for c in all_ctx)
    return ColorChainResult(stats_blend, tuple(float(f) for f in fac), tuple(bal), stf, l1 - l0, lut)


# ------------------------------------------------------------------------------------- process_fits
@dataclass
class ProcessFitsResult:
    stats: ImageStatsC
    stf: StfParamsC
    preview_u8: np.ndarray                 # rows x cols, what save_preview_png receives
    histogram_bins: Optional[np.ndarray]   # 512 display bins (process_fits_full), else None


def process_fits(unit, bitpix: int, bscale: float, bzero: float, rows: int, cols: int, full: bool = False,
                 ctx: Context | None = None) -> ProcessFitsResult:
    """The mono branch of process_fits / process_fits_full (cmd/io/mod.rs:23-31,103-170) between the
    file read and the PNG encoder, with the plane resident in HBM: the data unit is uploaded as bytes,
    decode_pixels (infra/fits/reader.rs:37-101), compute_image_stats, apply_stf and -- for the full
    command -- compute_histogram_with_stats run on the device; auto_stf and downsample_histogram are the
    scalar / 65,536-element host code.  One upload of the file's bytes, one download of the u8 preview."""
    import torch

    ctx = ctx or default_context()
    h = ctx.handle
    n = rows * cols
    raw = np.frombuffer(unit, np.uint8) if isinstance(unit, (bytes, bytearray, memoryview)) else np.ascontiguousarray(unit).view(np.uint8).reshape(-1)
    bpp = abs(int(bitpix)) // 8
    if bpp == 0 or raw.size < n * bpp:
        raise ValueError(f"data unit of {raw.size} bytes for {rows}x{cols} BITPIX {bitpix}")
    dev = torch.device("cuda", ctx.device)
    stream = torch.cuda.current_stream(dev).cuda_stream
    s = stream if stream else None
    d_raw = torch.from_numpy(raw[: n * bpp].copy()).to(dev, non_blocking=True)
    plane = torch.empty(n, dtype=torch.float32, device=dev)
    ctx.check(lib().ab_fits_decode_dev(h, d_raw.data_ptr(), n, int(bitpix), float(bscale), float(bzero), plane.data_ptr(), s))
    st = ImageStatsC()
    ctx.check(lib().ab_image_stats_dev(h, plane.data_ptr(), n, 0, 0.0, 0.0, C.addressof(st), s))
    cfg = AutoStfConfigC(0.25, -2.8)
    stf = StfParamsC()
    lib().ab_auto_stf(C.addressof(st), C.addressof(cfg), C.addressof(stf))
    u8 = torch.empty(n, dtype=torch.uint8, device=dev)
    ctx.check(lib().ab_apply_stf_u8_dev(h, plane.data_ptr(), n, C.addressof(stf), C.addressof(st), u8.data_ptr(), s))
    bins = None
    if full:
        hist = np.zeros(65536, np.uint32)
        edges = np.zeros(65537, np.float64)
        ctx.check(lib().ab_histogram_u32_dev(h, plane.data_ptr(), n, 65536, st.min, st.max, hist.ctypes.data,
                                             edges.ctypes.data, s))
        bins = np.zeros(512, np.uint32)
        ctx.check(lib().ab_downsample_histogram(hist.ctypes.data, hist.size, 512, bins.ctypes.data))
    return ProcessFitsResult(st, stf, u8.cpu().numpy().reshape(rows, cols), bins)


# ------------------------------------------------------------------------------------- config 3
def align_and_warp(ref_ptr: int, tgt_ptr: int, out_ptr: int, rows: int, cols: int, n_threads: int = 8,
                   ctx: Context | None = None, stream: int = 0) -> AlignResultC:
    """One target frame of align_channels_cmd (affine): align_channel_affine(ref, tgt) then
    warp_image(tgt, transform) into `out_ptr` (blend.rs:263-296)."""
    ctx = ctx or default_context()
    res = AlignResultC()
    s = stream if stream else None
    ctx.check(lib().ab_align_affine_dev(ctx.handle, ref_ptr, tgt_ptr, rows, cols, n_threads, C.addressof(res), s))
    ctx.check(lib().ab_warp_affine_bicubic_dev(ctx.handle, tgt_ptr, rows, cols, C.addressof(res.transform), rows, cols,
                                               out_ptr, s))
    return res


def align_ref_stars(ref_ptr: int, rows: int, cols: int, ctx: Context | None = None, stream: int = 0) -> np.ndarray:
    """The reference frame's star list as the aligner sees it (<= 120 (x, y) pairs, f64), computed
    once for all targets (ab_align_ref_stars_dev)."""
    ctx = ctx or default_context()
    xy = np.zeros((120, 2), np.float64)
    n = C.c_size_t(0)
    ctx.check(lib().ab_align_ref_stars_dev(ctx.handle, ref_ptr, rows, cols, xy.ctypes.data, C.byref(n),
                                           stream if stream else None))
    return xy[: n.value].copy()


def align_and_warp_from_stars(ref_xy: np.ndarray, tgt_xy: np.ndarray, ref_ptr: Optional[int], tgt_ptr: int,
                              out_ptr: int, rows: int, cols: int, n_threads: int = 8, ctx: Context | None = None,
                              stream: int = 0, warp_stream: Optional[int] = None) -> AlignResultC:
    """align_and_warp with both star lists given (`align_ref_stars` on each plane): the target's
    detection can run before the reference's list has arrived from the GPU that holds frame 0."""
    ctx = ctx or default_context()
    res = AlignResultC()
    s = stream if stream else None
    rxy = np.ascontiguousarray(ref_xy, dtype=np.float64)
    txy = np.ascontiguousarray(tgt_xy, dtype=np.float64)
    ctx.check(lib().ab_align_affine_from_stars_dev(ctx.handle, rxy.ctypes.data, len(rxy), txy.ctypes.data, len(txy),
                                                   ref_ptr, tgt_ptr, rows, cols, n_threads, C.addressof(res), s))
    ctx.check(lib().ab_warp_affine_bicubic_dev(ctx.handle, tgt_ptr, rows, cols, C.addressof(res.transform), rows, cols,
                                               out_ptr, warp_stream if warp_stream is not None else s))
    return res


def align_and_warp_with_ref(ref_xy: np.ndarray, ref_ptr: Optional[int], tgt_ptr: int, out_ptr: int, rows: int,
                            cols: int, n_threads: int = 8, ctx: Context | None = None, stream: int = 0,
                            warp_stream: Optional[int] = None) -> AlignResultC:
    """align_and_warp with the reference's star list given (broadcast to the GPU that holds the
    target); `ref_ptr` = None when the reference plane is not on this GPU.  The aligner returns
    after its last device read of the target, so the warp may go to another stream (`warp_stream`)
    and overlap the next frame's detection; the caller synchronises that stream before using `out`."""
    ctx = ctx or default_context()
    res = AlignResultC()
    s = stream if stream else None
    xy = np.ascontiguousarray(ref_xy, dtype=np.float64)
    ctx.check(lib().ab_align_affine_with_ref_dev(ctx.handle, xy.ctypes.data, len(xy), ref_ptr, tgt_ptr, rows, cols,
                                                 n_threads, C.addressof(res), s))
    ctx.check(lib().ab_warp_affine_bicubic_dev(ctx.handle, tgt_ptr, rows, cols, C.addressof(res.transform), rows, cols,
                                               out_ptr, warp_stream if warp_stream is not None else s))
    return res


def align_channels(ref_ptr: int, tgt_ptrs: Sequence[int], out_ptrs: Sequence[int], rows: int, cols: int,
                   n_threads: int = 8, lanes: int = 3, ctx: Context | None = None):
    """align_channels_cmd with alignMethod = "affine" (cmd/compose/blend.rs:263-296) for planes resident
    on one GPU: every target is registered against the reference and warped into its output plane.
    The reference's star list is computed once; `lanes` host threads (each with its own context and
    streams) feed the GPU, because the detector synchronises a few times per frame and another
    frame's kernels fill those gaps (16 x 8192^2 on a B200: 54 ms with one lane, 41 ms with three).
    Results are the per-pair ab_align_affine results, in target order; the call returns when every
    warp has finished."""
    import torch
    from concurrent.futures import ThreadPoolExecutor

    ctx = ctx or default_context()
    n = len(tgt_ptrs)
    if n == 0:
        return []
    lanes = max(1, min(int(lanes), n, 8))
    dev = torch.device("cuda", ctx.device)
    ctxs = [ctx] + [Context(ctx.device) for _ in range(lanes - 1)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)]
    sides = [torch.cuda.Stream(device=dev) for _ in range(lanes)]
    torch.cuda.synchronize(dev)      # whatever produced the planes has finished
    results = [None] * n
    try:
        with ThreadPoolExecutor(lanes) as pool:
            def detect_lane(k):
                return {i: align_ref_stars(tgt_ptrs[i], rows, cols, ctxs[k], streams[k].cuda_stream)
                        for i in range(k, n, lanes)}

            futs = [pool.submit(detect_lane, k) for k in range(1, lanes)]
            ref_xy = align_ref_stars(ref_ptr, rows, cols, ctx, streams[0].cuda_stream)
            tgt_xy = detect_lane(0)
            for fu in futs:
                tgt_xy.update(fu.result())

            def finish_lane(k):
                for i in range(k, n, lanes):
                    results[i] = align_and_warp_from_stars(ref_xy, tgt_xy[i], ref_ptr, tgt_ptrs[i], out_ptrs[i], rows,
                                                           cols, n_threads, ctxs[k], streams[k].cuda_stream,
                                                           warp_stream=sides[k].cuda_stream)

            futs = [pool.submit(finish_lane, k) for k in range(1, lanes)]
            finish_lane(0)
            for fu in futs:
                fu.result()
        torch.cuda.synchronize(dev)
    finally:
        for c in ctxs[1:]:
            c.close()
    return results
