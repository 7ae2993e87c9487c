#!/usr/bin/env python
"""bench_ops.py -- per-kernel roofline numbers for the non-headline rows of SURVEY 8(a) on one GPU
(BASELINE.json configs 1, 3, 4, 5).  Prints one JSON line per operator:

  {"op", "config", "ms", "algorithmic_bytes", "achieved_gbs", "frac_of_measured_peak",
   "mpix_s", "cpu_oracle_mpix_s" (bounded sample, cores stated), "parity" }

Device-resident inputs, CUDA events on the launching stream, best of `--reps` after warm-up;
planes are far larger than L2 (685 MB at config 4) or the L2 is flushed between repetitions.
bench.py stays the contract benchmark; this file is evidence for the other rows.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--small", action="store_true", help="quarter-size planes (quick check)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    import torch

    import astroburst_b200 as ab
    from astroburst_b200 import _lib, synth
    from astroburst_b200._lib import lib
    from oracle import oracle

    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    ctx = ab.default_context(0)
    h = ctx.handle
    peak = 6547.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    results = []

    def timed(fn, reps=args.reps):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            flush.zero_()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    def wall(fn, reps=3):
        fn()
        best = 1e30
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t0) * 1e3)
        return best

    def emit(op, config, ms, nbytes, npix, cpu=None, parity=None, note=None):
        line = {"op": op, "config": config, "ms": round(ms, 4), "algorithmic_bytes": int(nbytes),
                "achieved_gbs": round(nbytes / ms / 1e6, 1), "frac_of_measured_peak": round(nbytes / ms / 1e6 / peak, 3),
                "mpix_s": round(npix / ms / 1e3, 1)}
        if cpu is not None:
            line["cpu_oracle_mpix_s"] = round(cpu[0], 2)
            line["cpu_cores"] = cpu[1]
            line["cpu_sample"] = cpu[2]
        if parity is not None:
            line["parity"] = parity
        if note:
            line["note"] = note
        results.append(line)
        print(json.dumps(line), flush=True)

    def cpu_rate(fn, npix):
        fn()
        t0 = time.perf_counter()
        fn()
        return npix / (time.perf_counter() - t0) / 1e6, oracle.num_threads()

    P = C.c_void_p
    # ------------------------------------------------------------------ config 4: colour chain
    rows, cols = (13759, 12451) if not args.small else (6880, 6226)
    npix = rows * cols
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    yy = torch.arange(rows, device=dev, dtype=torch.float32)[:, None]
    xx = torch.arange(cols, device=dev, dtype=torch.float32)[None, :]
    chans = []
    for k in range(3):
        neb = 600.0 + 500.0 * torch.sin(yy * (0.0007 * (k + 1))) * torch.cos(xx * (0.0011 * (k + 2)))
        p = neb + torch.randn((rows, cols), generator=g, device=dev) * 30.0 + 200.0 * k
        p.clamp_(min=1.0)
        p[: rows // 40] = 0.0
        chans.append(p.contiguous())
    del yy, xx
    r = torch.empty((rows, cols), dtype=torch.float32, device=dev)
    gg = torch.empty_like(r)
    b = torch.empty_like(r)
    cfg4 = f"3 x {rows}x{cols} f32 (config 4), 1 GPU"
    w = (_lib.BlendWeightC * 3)(_lib.BlendWeightC(0, 0.8, 0.1, 0.0), _lib.BlendWeightC(1, 0.2, 0.7, 0.15),
                                _lib.BlendWeightC(2, 0.0, 0.2, 0.85))
    ptrs = (C.c_void_p * 3)(*[c.data_ptr() for c in chans])
    lens = (C.c_size_t * 3)(npix, npix, npix)
    ms = timed(lambda: ctx.check(lib().ab_blend_channels_dev(h, ptrs, lens, 3, w, 3, rows, cols, r.data_ptr(),
                                                             gg.data_ptr(), b.data_ptr(), stream)))
    band = slice(rows // 2, rows // 2 + 512)
    hb = [c[band].cpu().numpy() for c in chans]
    wl = [(0, 0.8, 0.1, 0.0), (1, 0.2, 0.7, 0.15), (2, 0.0, 0.2, 0.85)]
    ob = oracle.blend_channels(hb, wl, 512, cols)
    par = bool(np.array_equal(ob[0], r[band].cpu().numpy()) and np.array_equal(ob[2], b[band].cpu().numpy()))
    cr = cpu_rate(lambda: oracle.blend_channels(hb, wl, 512, cols), 512 * cols)
    emit("blend_channels (hubble_legacy 3x3)", cfg4, ms, 6 * npix * 4, npix, (cr[0], cr[1], "512 rows"), par)

    st = _lib.ImageStatsC()
    ms = wall(lambda: ctx.check(lib().ab_image_stats_dev(h, r.data_ptr(), npix, 0, 0.0, 0.0, C.addressof(st), stream)))
    sub = r[: 4200000 // cols + 1].cpu().numpy()
    so = oracle.compute_image_stats(sub)
    sg = _lib.ImageStatsC()
    ctx.check(lib().ab_image_stats_dev(h, r.data_ptr(), sub.size, 0, 0.0, 0.0, C.addressof(sg), stream))
    par = (sg.median, sg.mad, sg.min, sg.max, sg.valid_count) == (so.median, so.mad, so.min, so.max, so.valid_count)
    cr = cpu_rate(lambda: oracle.compute_image_stats(sub), sub.size)
    emit("compute_image_stats (65536-bin histogram path, 4 sweeps chained on the device)", cfg4, ms, 4 * npix * 4, npix,
         (cr[0], cr[1], f"{sub.shape[0]} rows"), par, "wall clock of one call: four sweeps + one-block walk kernels on one stream, one 56-byte D2H")

    ms = timed(lambda: ctx.check(lib().ab_scale_dev(h, r.data_ptr(), npix, 1.07, r.data_ptr(), stream)))
    emit("white-balance scale (v * factor, in place)", cfg4, ms, 2 * npix * 4, npix)
    scn = _lib.ScnrConfigC(0, 0.5, 0)
    ms = timed(lambda: ctx.check(lib().ab_apply_scnr_dev(h, r.data_ptr(), gg.data_ptr(), b.data_ptr(), npix,
                                                         C.addressof(scn), stream)))
    emit("apply_scnr_inplace (average, 0.5)", cfg4, ms, 4 * npix * 4, npix)
    scn2 = _lib.ScnrConfigC(0, 0.5, 1)
    ms = timed(lambda: ctx.check(lib().ab_apply_scnr_dev(h, r.data_ptr(), gg.data_ptr(), b.data_ptr(), npix,
                                                         C.addressof(scn2), stream)))
    emit("apply_scnr_inplace (preserve luminance)", cfg4, ms, 6 * npix * 4, npix)

    ctx.check(lib().ab_image_stats_dev(h, gg.data_ptr(), npix, 0, 0.0, 0.0, C.addressof(st), stream))
    cfgs = _lib.AutoStfConfigC(0.25, -2.8)
    stf = _lib.StfParamsC()
    lib().ab_auto_stf(C.addressof(st), C.addressof(cfgs), C.addressof(stf))
    out = torch.empty_like(r)
    ms = timed(lambda: ctx.check(lib().ab_apply_stf_f32_dev(h, gg.data_ptr(), npix, C.addressof(stf), C.addressof(st), 0,
                                                            out.data_ptr(), stream)))
    hb1 = gg[band].cpu().numpy()
    ost = oracle.ImageStats(st.min, st.max, st.median, st.mad, st.sigma, st.mean, st.valid_count)
    osf = oracle.StfParams(stf.shadow, stf.midtone, stf.highlight)
    oo = oracle.apply_stf_f32(hb1, osf, ost)
    got = out[band].cpu().numpy()
    par = bool(np.allclose(got, oo, rtol=1e-5, atol=0))
    cr = cpu_rate(lambda: oracle.apply_stf_f32(hb1, osf, ost), hb1.size)
    emit("apply_stf_f32 (linked auto-STF)", cfg4, ms, 2 * npix * 4, npix, (cr[0], cr[1], "512 rows"), par)
    u8 = torch.empty((rows, cols), dtype=torch.uint8, device=dev)
    ms = timed(lambda: ctx.check(lib().ab_apply_stf_u8_dev(h, gg.data_ptr(), npix, C.addressof(stf), C.addressof(st),
                                                           u8.data_ptr(), stream)))
    par = bool(np.array_equal(u8[band].cpu().numpy(), oracle.apply_stf(hb1, osf, ost)))
    emit("apply_stf (u8 preview)", cfg4, ms, npix * 5, npix, None, par)
    lut = torch.from_numpy(ab.imaging.spline_lut([(0, 0), (0.25, 0.15), (0.5, 0.5), (0.75, 0.85), (1, 1)])).to(dev)
    ms = timed(lambda: ctx.check(lib().ab_apply_curve_dev(h, out.data_ptr(), npix, lut.data_ptr(), r.data_ptr(), stream)))
    emit("apply_curve (4096-entry LUT)", cfg4, ms, 2 * npix * 4, npix)
    lv = _lib.LevelsParamsC(0.02, 1.8, 0.97)
    ms = timed(lambda: ctx.check(lib().ab_apply_levels_dev(h, out.data_ptr(), npix, C.addressof(lv), r.data_ptr(), stream)))
    emit("apply_levels (f64 powf)", cfg4, ms, 2 * npix * 4, npix)
    ms = timed(lambda: ctx.check(lib().ab_arcsinh_stretch_dev(h, gg.data_ptr(), npix, 1.0, 3000.0, 50.0, 1.0,
                                                              out.data_ptr(), stream)))
    emit("arcsinh_stretch (factor 50)", cfg4, ms, 2 * npix * 4, npix)
    # fused tone chain: linked STF -> curves -> SCNR in one sweep (3 planes in, 3 planes out)
    tc = _lib.ToneChainC()
    for k in range(3):
        tc.stf[k], tc.stats[k] = stf, st
        tc.levels[k] = _lib.LevelsParamsC(0.0, 1.0, 1.0)
        tc.lut[k] = lut.data_ptr()
    tc.apply_scnr, tc.scnr = 1, _lib.ScnrConfigC(0, 0.5, 0)
    o3 = [torch.empty_like(gg) for _ in range(3)]
    src3 = [chans[0], gg, chans[2]]
    ms = timed(lambda: ctx.check(lib().ab_tone_composite_dev(h, src3[0].data_ptr(), src3[1].data_ptr(), src3[2].data_ptr(),
                                                             npix, C.addressof(tc), o3[0].data_ptr(), o3[1].data_ptr(),
                                                             o3[2].data_ptr(), stream)))
    hb3 = [c[band].cpu().numpy() for c in src3]
    w3 = [oracle.apply_curve(oracle.apply_stf_f32(x, osf, ost), lut.cpu().numpy()) for x in hb3]
    oracle.apply_scnr_inplace(*w3, "average", 0.5, False)
    par = all(np.array_equal(o[band].cpu().numpy(), w) for o, w in zip(o3, w3))
    emit("tone chain fused (STF + curve + SCNR, one sweep)", cfg4, ms, 6 * npix * 4, npix, None, par,
         "the four separate operators move 20 planes; the fused sweep moves 6")
    # config-4 chain of SURVEY 8(d) from the measured stages: blend + 3 plane stats + WB (3 scales)
    # + SCNR + G re-stat + fused tone = 20.56 GB algorithmic on one GPU
    def stage(prefix):
        return next(x["ms"] for x in results if x["op"].startswith(prefix))

    chain_ms = (stage("blend_channels") + 4 * stage("compute_image_stats") + 3 * stage("white-balance scale")
                + stage("apply_scnr_inplace (average") + stage("tone chain fused"))
    emit("config-4 colour chain (sum of the measured stages: blend, 3+1 stats, WB, SCNR, fused tone)", cfg4,
         chain_ms, int(20.56e9), npix, None, None, "SURVEY 8(d): 20.56 GB algorithmic, ideal 3.14 ms at the measured peak")
    # RGB8 interleave: full size and the 2048-px preview (cmd/helpers.rs MAX_PREVIEW_DIM)
    rgb8 = torch.empty((rows, cols, 3), dtype=torch.uint8, device=dev)
    ms = timed(lambda: ctx.check(lib().ab_render_rgb_u8_dev(h, o3[0].data_ptr(), o3[1].data_ptr(), o3[2].data_ptr(),
                                                            rows, cols, 0, rgb8.data_ptr(), stream)))
    par = bool(np.array_equal(rgb8[band].cpu().numpy(),
                              oracle.render_rgb_u8(*[o[band].cpu().numpy() for o in o3])))
    emit("render_rgb (RGB8 interleave, full size)", cfg4, ms, npix * 15, npix, None, par)
    ms = timed(lambda: ctx.check(lib().ab_render_rgb_u8_dev(h, o3[0].data_ptr(), o3[1].data_ptr(), o3[2].data_ptr(),
                                                            rows, cols, 2048, rgb8.data_ptr(), stream)))
    pw, ph = ab.preview.preview_size(rows, cols, 2048)
    emit("render_rgb_preview (nearest decimation to 2048 px)", cfg4, ms, pw * ph * 15, pw * ph, None, None,
         "gather: one sector per sampled pixel")
    del chans, r, gg, b, out, u8, o3, rgb8, src3
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ masked stretch (N4) on a 4096^2 star field
    Rm = 4096 if not args.small else 2048
    fm = synth.stack_frames_torch(1, Rm, Rm, dev, seed=9, nan_border=0, cosmic_fraction=0.0)[0]
    fm = ((fm - 900.0).clamp_(min=0.0) / 20000.0).contiguous()
    msk = torch.empty_like(fm)
    mo = torch.empty_like(fm)
    mcfg = _lib.MaskedStretchConfigC(10, 0.25, 2.5, 4.0, 1, 0.85, 0.85, 1e-5)
    minfo = _lib.MaskedStretchInfoC()
    ms = wall(lambda: ctx.check(lib().ab_masked_stretch_dev(h, fm.data_ptr(), Rm, Rm, C.addressof(mcfg), msk.data_ptr(),
                                                            mo.data_ptr(), C.addressof(minfo), stream)))
    subm = fm[:1024, :1024].contiguous()
    o_img, o_it, o_fb, o_conv, o_n, o_cov = oracle.masked_stretch(subm.cpu().numpy())
    mi2 = _lib.MaskedStretchInfoC()
    so, sk = torch.empty_like(subm), torch.empty_like(subm)
    ctx.check(lib().ab_masked_stretch_dev(h, subm.data_ptr(), 1024, 1024, C.addressof(mcfg), sk.data_ptr(), so.data_ptr(),
                                          C.addressof(mi2), stream))
    par = bool(np.array_equal(so.cpu().numpy(), o_img)) and (int(mi2.iterations_run), int(mi2.stars_masked)) == (o_it, o_n)
    cr = cpu_rate(lambda: oracle.masked_stretch(subm.cpu().numpy()), subm.numel())
    itn = int(minfo.iterations_run)
    emit(f"masked_stretch (detect + mask + {itn} iterations, {int(minfo.stars_masked)} stars)", f"{Rm}x{Rm} f32", ms,
         Rm * Rm * 4 * (4 + itn * (4 * 2 + 3) + 4 * 2 + 2), Rm * Rm, (cr[0], cr[1], "1024x1024 crop"), par,
         "wall clock: detection + mask + per iteration 4 radix sweeps (pixel + mask) and one fused MTF/blend sweep")
    del fm, msk, mo
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ config 3: detect + align + warp
    R3 = 8192 if not args.small else 4096
    cfg3 = f"{R3}x{R3} f32 frame (config 3), 1 GPU"
    rng = np.random.default_rng(3)
    nst = 3000 * (R3 // 4096) ** 2
    x, y = rng.uniform(20, R3 - 20, nst), rng.uniform(20, R3 - 20, nst)
    fl = 300.0 * (40000.0 / 300.0) ** (rng.uniform(0, 1, nst) ** 2)

    def frame(t, seed):
        g3 = torch.Generator(device=dev)
        g3.manual_seed(seed)
        img = torch.randn((R3, R3), generator=g3, device=dev) * 4.0 + 200.0
        xs = torch.tensor(t[0] * x + t[1] * y + t[2], device=dev)
        ys = torch.tensor(t[3] * x + t[4] * y + t[5], device=dev)
        f = torch.tensor(fl, device=dev)
        rr = 7
        dy, dx = torch.meshgrid(torch.arange(-rr, rr + 1, device=dev), torch.arange(-rr, rr + 1, device=dev), indexing="ij")
        cx, cy = xs.round().long(), ys.round().long()
        px, py = cx[:, None, None] + dx[None], cy[:, None, None] + dy[None]
        val = f[:, None, None] / (2 * math.pi * 1.5 ** 2) * torch.exp(
            -((px - xs[:, None, None]) ** 2 + (py - ys[:, None, None]) ** 2) / (2 * 1.5 ** 2))
        ok = (px >= 0) & (px < R3) & (py >= 0) & (py < R3)
        img.view(-1).index_add_(0, (py * R3 + px)[ok], val[ok].float())
        return img.contiguous()

    th = math.radians(1.1)
    t_true = (1.004 * math.cos(th), -math.sin(th), 23.5, math.sin(th), 1.004 * math.cos(th), -17.25)
    ref = frame((1, 0, 0, 0, 1, 0), 1)
    tgt = frame(t_true, 2)
    n3 = R3 * R3
    info = _lib.DetectionInfoC()
    stars = (_lib.StarC * 100000)()
    ms = wall(lambda: ctx.check(lib().ab_detect_stars_dev(h, ref.data_ptr(), R3, R3, 3.5, stars, 100000,
                                                          C.addressof(info), stream)))
    sub = ref[:2048, :2048].contiguous()
    subn = sub.cpu().numpy()
    od = oracle.detect_stars(subn, 3.5)
    info2 = _lib.DetectionInfoC()
    ctx.check(lib().ab_detect_stars_dev(h, sub.data_ptr(), 2048, 2048, 3.5, stars, 100000, C.addressof(info2), stream))
    par = int(info2.n_stars) == len(od.stars) and info2.background_median == od.background_median
    cr = cpu_rate(lambda: oracle.detect_stars(subn, 3.5), subn.size)
    emit(f"detect_stars ({int(info.n_stars)} stars)", cfg3, ms, 2 * n3 * 4, n3, (cr[0], cr[1], "2048x2048 crop"), par,
         "wall clock: 2 GPU sweeps + tile selects + host flood fill of the sparse list")
    res = _lib.AlignResultC()
    ms = wall(lambda: ctx.check(lib().ab_align_affine_dev(h, ref.data_ptr(), tgt.data_ptr(), R3, R3, 16,
                                                          C.addressof(res), stream)), reps=2)
    got = [getattr(res.transform, k) for k in "a b tx c d ty".split()]
    err = max(math.hypot(got[0] * X + got[1] * Y + got[2] - (t_true[0] * X + t_true[1] * Y + t_true[2]),
                         got[3] * X + got[4] * Y + got[5] - (t_true[3] * X + t_true[4] * Y + t_true[5]))
              for X, Y in ((0, 0), (R3, 0), (0, R3), (R3, R3)))
    emit(f"align_channel_affine (method {res.method}, {res.inliers} inliers, corner error {err:.3f} px)", cfg3, ms,
         4 * n3 * 4, n3, None, bool(err < 1.0), "wall clock: 2 x (normalise + detect) + triangle votes + host RANSAC")
    wout = torch.empty_like(tgt)
    ms = timed(lambda: ctx.check(lib().ab_warp_affine_bicubic_dev(h, tgt.data_ptr(), R3, R3, C.addressof(res.transform),
                                                                  R3, R3, wout.data_ptr(), stream)))
    tn = tgt[:1024].cpu().numpy()
    ow = oracle.warp_image(tgt.cpu().numpy(), tuple(got), 256, R3)
    par = bool(np.array_equal(ow, wout[:256].cpu().numpy()))
    big = tgt.cpu().numpy()
    cr = cpu_rate(lambda: oracle.warp_image(big, tuple(got), 512, R3), 512 * R3)
    emit("warp_image (bicubic Catmull-Rom)", cfg3, ms, 2 * n3 * 4, n3, (cr[0], cr[1], "512 output rows"), par)
    del ref, tgt, wout, big
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ config 1: 1600^2 auto-STF
    img = synth.stack_frames(1, 1600, 1600, seed=9, nan_border=0)[0]
    img[:60] = 0.0
    d1 = torch.from_numpy(img).to(dev)
    st1 = _lib.ImageStatsC()
    u1 = torch.empty((1600, 1600), dtype=torch.uint8, device=dev)

    def c1():
        ctx.check(lib().ab_image_stats_dev(h, d1.data_ptr(), d1.numel(), 0, 0.0, 0.0, C.addressof(st1), stream))
        lib().ab_auto_stf(C.addressof(st1), C.addressof(cfgs), C.addressof(stf))
        ctx.check(lib().ab_apply_stf_u8_dev(h, d1.data_ptr(), d1.numel(), C.addressof(stf), C.addressof(st1),
                                            u1.data_ptr(), stream))

    ms = wall(c1)
    o1 = oracle.compute_image_stats(img)
    op = oracle.auto_stf(o1)
    par = bool((st1.median, st1.mad) == (o1.median, o1.mad) and np.array_equal(oracle.apply_stf(img, op, o1), u1.cpu().numpy()))

    def c1cpu():
        s_ = oracle.compute_image_stats(img)
        oracle.apply_stf(img, oracle.auto_stf(s_), s_)

    cr = cpu_rate(c1cpu, img.size)
    emit("stats (exact select) + auto_stf + apply_stf u8", "1600x1600 f32 (config 1)", ms, 3 * img.size * 4, img.size,
         (cr[0], cr[1], "full frame"), par, "wall clock: scan + <= 16 radix passes (host-stepped) + STF sweep")

    # ------------------------------------------------------------------ config 5: 256-frame deep stack
    rows5 = 512 if not args.small else 64  # 512 rows = the band one of 8 GPUs owns
    fr = synth.stack_frames_torch(256, rows5, 4096, dev, seed=5, nan_border=0)
    cube = torch.stack(fr)
    del fr
    o5 = torch.empty((rows5, 4096), dtype=torch.float32, device=dev)
    rej5 = torch.zeros(1, dtype=torch.int64, device=dev)
    cfg5 = ab.StackConfig(3.0, 3.0, 5, False)
    p5 = [cube[f].data_ptr() for f in range(256)]
    ms = timed(lambda: ab.stack_images_device(p5, rows5 * 4096, cfg5, o5.data_ptr(), rej5.data_ptr(), stream, ctx), reps=2)
    band5 = [cube[f, :8].cpu().numpy() for f in range(256)]
    w5, _ = oracle.stack_images(band5, 3.0, 3.0, 5)
    par = bool(np.array_equal(w5, o5[:8].cpu().numpy()))
    cr = cpu_rate(lambda: oracle.stack_images(band5, 3.0, 3.0, 5), 8 * 4096)
    emit("256-frame sigma-clip stack (chunked deep kernel)", f"256 x {rows5}x4096 f32 row band (config 5: what one of 8 GPUs stacks)",
         ms, 257 * rows5 * 4096 * 4, rows5 * 4096, (cr[0], cr[1], "8 rows"), par)

    if args.out:
        with open(args.out, "w") as f:
            json.dump({"peak_gbs": peak, "results": results}, f, indent=1)


if __name__ == "__main__":
    main()
