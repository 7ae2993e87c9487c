"""The command-order compositions (astroburst_b200/pipelines.py) on the GPU against the same
compositions of the CPU oracle (oracle/chains.py): config-4 colour chain (exact-statistics and
histogram-statistics sizes) and the config-3 align + warp step with a shared reference star list."""
import ctypes as C
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WEIGHTS = [(0, 0.8, 0.1, 0.0), (1, 0.2, 0.7, 0.15), (2, 0.0, 0.2, 0.85)]
CURVE = [(0.0, 0.0), (0.25, 0.15), (0.5, 0.5), (0.75, 0.85), (1.0, 1.0)]


def _chans(rows, cols, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:rows, 0:cols].astype(np.float32)
    out = []
    for k in range(3):
        p = 600.0 + 500.0 * np.sin(yy * 0.01 * (k + 1)) * np.cos(xx * 0.013 * (k + 2))
        p = (p + rng.normal(0, 30.0, (rows, cols)) + 200.0 * k).astype(np.float32)
        np.maximum(p, 1.0, out=p)
        p[: rows // 20] = 0.0
        out.append(p)
    return out


@pytest.mark.parametrize("rows,cols,scnr", [(300, 420, (0, 0.5, 0)), (300, 420, (1, 0.8, 1)), (2100, 2000, (0, 0.5, 0)),
                                            (2100, 2000, (0, 0.0, 0))])
def test_color_chain_matches_oracle_chain(orc, rows, cols, scnr):
    import torch
    import astroburst_b200 as ab
    from astroburst_b200 import pipelines
    from oracle import chains

    dev = torch.device("cuda:0")
    ctx = ab.default_context(0)
    stream = torch.cuda.current_stream().cuda_stream
    ch = _chans(rows, cols, rows)
    d = [torch.from_numpy(c).to(dev) for c in ch]
    out = [torch.empty((rows, cols), dtype=torch.float32, device=dev) for _ in range(3)]
    res = pipelines.color_chain([x.data_ptr() for x in d], rows * cols, rows, cols, WEIGHTS, CURVE,
                                out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                pipelines.stats_single_gpu(ctx, stream), scnr_after_wb=scnr, ctx=ctx, stream=stream)
    torch.cuda.synchronize()
    want, sb, bal, stf = chains.color_chain(ch, WEIGHTS, CURVE, (("average", "maximum")[scnr[0]], scnr[1], bool(scnr[2])))
    key = lambda s: (s.min, s.max, s.median, s.mad, s.sigma, int(s.valid_count))
    assert [key(s) for s in res.stats_blend] == [key(s) for s in sb]
    assert [key(s) for s in res.stats_balanced] == [key(s) for s in bal]
    assert (res.stf.shadow, res.stf.midtone, res.stf.highlight) == (stf.shadow, stf.midtone, stf.highlight)
    for g, w in zip(out, want):
        g = g.cpu().numpy()
        bad = ~(np.isclose(g, w, rtol=1e-5, atol=0.0) | (g == w))     # north_star: 1e-5 relative
        assert bad.mean() < 1e-3                                      # (a 1-ulp STF difference may pick the next LUT bin)
        assert np.abs(g - w).max() < 2e-3


def test_align_with_shared_reference_star_list_equals_pairwise_align(orc):
    import importlib, os, sys
    import torch
    import astroburst_b200 as ab
    from astroburst_b200 import pipelines
    from astroburst_b200._lib import AlignResultC, lib

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    star_field = importlib.import_module("test_stars_gpu").star_field
    rows = cols = 1024
    th = math.radians(0.8)
    t_true = (1.002 * math.cos(th), -math.sin(th), 9.5, math.sin(th), 1.002 * math.cos(th), -6.25)
    ref = star_field(rows, cols, 250, seed=11)
    tgts = [star_field(rows, cols, 250, seed=11, transform=t_true, noise_seed=k) for k in (1, 2)]
    dev = torch.device("cuda:0")
    ctx = ab.default_context(0)
    s = torch.cuda.current_stream().cuda_stream
    d_ref = torch.from_numpy(ref).to(dev)
    xy = pipelines.align_ref_stars(d_ref.data_ptr(), rows, cols, ctx, s)
    assert 4 <= len(xy) <= 120
    for t in tgts:
        d_t = torch.from_numpy(t).to(dev)
        out = torch.empty_like(d_t)
        got = pipelines.align_and_warp_with_ref(xy, d_ref.data_ptr(), d_t.data_ptr(), out.data_ptr(), rows, cols, 8, ctx, s)
        pair = AlignResultC()
        ctx.check(lib().ab_align_affine_dev(ctx.handle, d_ref.data_ptr(), d_t.data_ptr(), rows, cols, 8, C.addressof(pair), s))
        tup = lambda r: tuple(getattr(r.transform, k) for k in "a b tx c d ty".split()) + (r.method, r.inliers, r.residual_px)
        assert tup(got) == tup(pair)
        o = orc.align_channel_affine(ref, t, 8)
        assert tup(got)[:6] == o.transform.as_tuple() and got.method == o.method
        torch.cuda.synchronize()
        want = orc.warp_image(t, tup(got)[:6], rows, cols)
        assert np.array_equal(out.cpu().numpy(), want)
        # without the reference plane on this GPU only the phase-correlation fallback is lost
        got2 = pipelines.align_and_warp_with_ref(xy, None, d_t.data_ptr(), out.data_ptr(), rows, cols, 8, ctx, s)
        assert tup(got2) == tup(got)
        # both star lists detected up front (the target's before the reference's list "arrives")
        txy = pipelines.align_ref_stars(d_t.data_ptr(), rows, cols, ctx, s)
        out3 = torch.empty_like(d_t)
        got3 = pipelines.align_and_warp_from_stars(xy, txy, d_ref.data_ptr(), d_t.data_ptr(), out3.data_ptr(), rows, cols, 8, ctx, s)
        torch.cuda.synchronize()
        assert tup(got3) == tup(pair) and np.array_equal(out3.cpu().numpy(), want)
    # too few stars on one side: the phase-correlation fallback, exactly as the pairwise call
    flat = torch.full((rows, cols), 0.25, device=dev)
    fxy = pipelines.align_ref_stars(flat.data_ptr(), rows, cols, ctx, s)
    assert len(fxy) < 4
    out4 = torch.empty_like(flat)
    got4 = pipelines.align_and_warp_from_stars(xy, fxy, d_ref.data_ptr(), flat.data_ptr(), out4.data_ptr(), rows, cols, 8, ctx, s)
    pair4 = AlignResultC()
    ctx.check(lib().ab_align_affine_dev(ctx.handle, d_ref.data_ptr(), flat.data_ptr(), rows, cols, 8, C.addressof(pair4), s))
    assert tup(got4) == tup(pair4) and got4.method in (2, 3)


@pytest.mark.parametrize("bitpix,shape", [(16, (700, 900)), (-32, (2100, 2000)), (16, (2100, 2000)), (8, (300, 301)), (-64, (257, 129))])
def test_process_fits_chain_matches_oracle_chain(orc, bitpix, shape):
    """pipelines.process_fits (mono branch of process_fits_full, cmd/io/mod.rs:23-31,103-170) against the
    same composition of the oracle: statistics (exact-select and histogram sizes), auto-STF, the u8
    preview plane and the 512 display bins, bit for bit."""
    from astroburst_b200 import pipelines
    from oracle import chains

    rows, cols = shape
    rng = np.random.default_rng(abs(bitpix) + rows)
    sky = 900.0 + rng.normal(0.0, 15.0, shape)
    sky[rng.random(shape) < 1e-3] += rng.uniform(500, 30000)
    sky[: rows // 25] = 0.0                                   # padding strip (invalid: <= 1e-7)
    if bitpix == 16:
        unit, bscale, bzero = np.clip(np.rint(sky - 32768.0), -32768, 32767).astype(">i2").tobytes(), 1.0, 32768.0
    elif bitpix == 8:
        unit, bscale, bzero = np.clip(np.rint(sky / 128.0), 0, 255).astype(np.uint8).tobytes(), 128.0, 0.0
    elif bitpix == -32:
        unit, bscale, bzero = sky.astype(">f4").tobytes(), 1.0, 0.0
    else:
        unit, bscale, bzero = sky.astype(">f8").tobytes(), 0.5, 3.0
    got = pipelines.process_fits(unit, bitpix, bscale, bzero, rows, cols, full=True)
    stats, stf, u8, bins = chains.process_fits(unit, bitpix, bscale, bzero, rows, cols, full=True)
    key = lambda s: (s.min, s.max, s.median, s.mad, s.sigma, int(s.valid_count))
    assert key(got.stats) == key(stats) and abs(got.stats.mean - stats.mean) <= 1e-12 * abs(stats.mean)
    assert (got.stf.shadow, got.stf.midtone, got.stf.highlight) == (stf.shadow, stf.midtone, stf.highlight)
    assert np.array_equal(got.preview_u8, np.asarray(u8).reshape(rows, cols))
    assert np.array_equal(got.histogram_bins, bins)


def test_align_channels_with_lanes_equals_pairwise_align(orc):
    """pipelines.align_channels (several host threads / contexts feeding one GPU) returns, for every
    target, exactly what the per-pair aligner returns, and the warped planes equal the oracle's."""
    import importlib, os, sys
    import torch
    import astroburst_b200 as ab
    from astroburst_b200 import pipelines
    from astroburst_b200._lib import AlignResultC, lib

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    star_field = importlib.import_module("test_stars_gpu").star_field
    rows = cols = 768
    dev = torch.device("cuda:0")
    ctx = ab.default_context(0)
    s = torch.cuda.current_stream().cuda_stream
    ref = star_field(rows, cols, 220, seed=21)
    tfs = []
    for k in range(5):
        th = math.radians(0.3 * (k + 1))
        tfs.append((math.cos(th), -math.sin(th), 3.5 * k - 4.0, math.sin(th), math.cos(th), 2.25 - 1.5 * k))
    tgts = [star_field(rows, cols, 220, seed=21, transform=t, noise_seed=40 + k) for k, t in enumerate(tfs)]
    tgts.append(np.full((rows, cols), 0.5, np.float32))            # no stars: the fallback path inside a lane
    d_ref = torch.from_numpy(ref).to(dev)
    d_t = [torch.from_numpy(t).to(dev) for t in tgts]
    outs = [torch.empty_like(x) for x in d_t]
    tup = lambda r: tuple(getattr(r.transform, k) for k in "a b tx c d ty".split()) + (r.method, r.inliers, r.residual_px)
    for lanes in (1, 3):
        res = pipelines.align_channels(d_ref.data_ptr(), [x.data_ptr() for x in d_t], [o.data_ptr() for o in outs],
                                       rows, cols, 8, lanes, ctx)
        for k, (r, x, o) in enumerate(zip(res, d_t, outs)):
            pair = AlignResultC()
            ctx.check(lib().ab_align_affine_dev(ctx.handle, d_ref.data_ptr(), x.data_ptr(), rows, cols, 8, C.addressof(pair), s))
            assert tup(r) == tup(pair), (lanes, k)
            want = orc.warp_image(tgts[k], tup(r)[:6], rows, cols)
            assert np.array_equal(o.cpu().numpy(), want), (lanes, k)
