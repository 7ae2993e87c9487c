"""CPU compositions of oracle functions in the order the reference's commands fix -- the checker of
astroburst_b200/pipelines.py.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.h): imported by tests/
and bench.py's checks, never by the product package."""
from __future__ import annotations

import numpy as np

from . import oracle

PAR_THRESHOLD = 4_000_000  # cmd/compose/color.rs:19


def auto_wb_factors(sr, sg, sb):
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


def color_chain(chans, weights, curve_points, scnr_after_wb=("average", 0.5, False),
                       stats_blend=None, stats_balanced=None):
    """blend_channels_cmd -> compute_auto_wb_cmd -> calibrate_and_scnr_cmd -> apply_tone_composite_cmd
    (linked STF + curve), cmd/compose/{blend,color}.rs, cmd/processing/curves.rs.  With `stats_blend` /
    `stats_balanced` given (tuples of oracle.ImageStats) the statistics are taken as is, so a row
    band can be checked against the banded GPU chain whose statistics are those of the whole image."""
    rows, cols = chans[0].shape
    r, g, b = oracle.blend_channels(chans, weights, rows, cols)
    if stats_blend is None:
        stats_blend = tuple(oracle.compute_image_stats(x) for x in (r, g, b))
    fac = [max(np.float32(f), np.float32(1e-6)) for f in auto_wb_factors(*stats_blend)]
    r, g, b = [oracle.scale(x, float(f)) for x, f in zip((r, g, b), fac)]
    if stats_balanced is None:
        npx = rows * cols
        bal = []
        for x, f, st in zip((r, g, b), fac, stats_blend):
            bal.append(oracle.compute_image_stats(x) if npx <= PAR_THRESHOLD else
                       oracle.compute_image_stats_with_known_range(x, st.min * float(f), st.max * float(f)))
    method, amount, preserve = scnr_after_wb
    if amount > 1e-7:
        oracle.apply_scnr_inplace(r, g, b, method, amount, preserve)
        if stats_balanced is None:
            if preserve:
                bal = [oracle.compute_image_stats(x) for x in (r, g, b)]
            else:
                bal[1] = oracle.compute_image_stats(g)
    if stats_balanced is not None:
        bal = list(stats_balanced)
    combined = oracle.linked_stats(*bal)
    stf = oracle.auto_stf(combined)
    lut = oracle.spline_lut(curve_points)
    out = [oracle.apply_curve(oracle.apply_stf_f32(x, stf, combined), lut) for x in (r, g, b)]
    return out, stats_blend, tuple(bal), stf




def process_fits(unit: bytes, bitpix: int, bscale: float, bzero: float, rows: int, cols: int, full: bool = False):
    """process_fits / process_fits_full for a mono image, cmd/io/mod.rs:23-31,103-170 (after the file
    read, before the PNG encoder): decode_pixels -> compute_image_stats -> auto_stf(default) ->
    apply_stf (u8) [-> compute_histogram_with_stats -> downsample_histogram(512)]."""
    img = oracle.fits_decode(unit, bitpix, bscale, bzero).reshape(rows, cols)
    stats = oracle.compute_image_stats(img)
    stf = oracle.auto_stf(stats)
    u8 = oracle.apply_stf(img, stf, stats)
    bins = None
    if full:
        hist, _ = oracle.build_histogram(img, 65536, stats.min, stats.max)
        bins = oracle.downsample_histogram(hist, 512)
    return stats, stf, u8, bins
