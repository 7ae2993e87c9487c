#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference's own sample data (run in the build container only:
/root/reference does not exist on the GPU box).

    python tests/golden/make_golden.py

Inputs : /root/reference/exampleFits/sample-data/{502,656,673}nmos.fits (HST WFPC2 M16, 1600x1600,
         BITPIX -32) -- the real fixtures SURVEY 8(c) names for BASELINE config 1.  A 192x192 crop of
         each (rows 600.., cols 120..: nebula, stars, exact-zero mosaic padding, negative pixels) is
         stored as the raw big-endian FITS data-unit bytes, so the FITS decoder is part of the path.
Outputs: m16_golden.npz -- what the CPU oracle computes on those crops: ImageStats, auto-STF
         parameters, STF u8 previews, star lists, and the linked-STF -> SCNR -> RGB8 compose tail.
The Rust reference cannot run here (no toolchain), so these vectors pin the ORACLE on real data
(regression fixtures); they are not outputs of the reference binary.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle  # noqa: E402

SRC = "/root/reference/exampleFits/sample-data"
R0, C0, N = 600, 120, 192
CHANNELS = ("502nmos", "656nmos", "673nmos")  # OIII, Halpha, SII


def data_unit(path):
    raw = open(path, "rb").read()
    cards = [raw[i:i + 80] for i in range(0, 2880 * 40, 80)]
    end = next(i for i, c in enumerate(cards) if c.startswith(b"END"))
    off = ((end + 1) * 80 + 2879) // 2880 * 2880
    return np.frombuffer(raw[off:off + 1600 * 1600 * 4], ">f4").reshape(1600, 1600)


def stats_vec(s):
    return np.array([s.min, s.max, s.median, s.mad, s.sigma, s.mean, float(s.valid_count)], np.float64)


def compute(crops):
    """Everything the goldens hold, from three float32 crops (also used by the tests)."""
    out = {}
    st = [oracle.compute_image_stats(c) for c in crops]
    stf = [oracle.auto_stf(s) for s in st]
    for k, name in enumerate(CHANNELS):
        out[f"stats_{name}"] = stats_vec(st[k])
        out[f"stf_{name}"] = np.array([stf[k].shadow, stf[k].midtone, stf[k].highlight])
        out[f"u8_{name}"] = oracle.apply_stf(crops[k], stf[k], st[k])
        det = oracle.detect_stars(crops[k], 5.0)
        out[f"stars_{name}"] = np.array([[s["x"], s["y"], s["flux"], s["fwhm"], s["eccentricity"], s["peak"],
                                          float(s["npix"]), s["snr"]] for s in det.stars], np.float64).reshape(-1, 8)
        out[f"bg_{name}"] = np.array([det.background_median, det.background_sigma])
    # compose tail (SHO palette): linked STF -> f32 stretch -> SCNR(average, 0.5) -> RGB8
    comb = oracle.linked_stats(st[2], st[1], st[0])
    lstf = oracle.auto_stf(comb)
    out["linked_stats"] = stats_vec(comb)
    out["linked_stf"] = np.array([lstf.shadow, lstf.midtone, lstf.highlight])
    rgb = [np.ascontiguousarray(oracle.apply_stf_f32(crops[k], lstf, comb)) for k in (2, 1, 0)]
    oracle.apply_scnr_inplace(*rgb, "average", 0.5, False)
    out["rgb8"] = oracle.render_rgb_u8(*rgb)
    return out


def load_crops():
    crops = []
    for name in CHANNELS:
        raw = open(os.path.join(HERE, f"m16_{name}_crop{N}.bef32"), "rb").read()
        crops.append(oracle.fits_decode(raw, -32).reshape(N, N))
    return crops


def main():
    for name in CHANNELS:
        d = data_unit(os.path.join(SRC, name + ".fits"))[R0:R0 + N, C0:C0 + N]
        open(os.path.join(HERE, f"m16_{name}_crop{N}.bef32"), "wb").write(np.ascontiguousarray(d).tobytes())
    crops = load_crops()
    for c in crops:
        print("crop: zeros", int((c == 0).sum()), "negative", int((c < 0).sum()), "max", float(c.max()))
    g = compute(crops)
    np.savez_compressed(os.path.join(HERE, "m16_golden.npz"), **g)
    print({k: (v.shape if v.ndim else float(v)) for k, v in g.items()})


if __name__ == "__main__":
    main()
