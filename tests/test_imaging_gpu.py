"""GPU parity tests for stats / histogram / STF / stretch / blend / SCNR / levels / curves and the
bicubic gathers: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): integer results (histogram bins, valid_count, u8 STF output)
bit-exact; f32 pixels within 1e-5 relative (most are bit-exact, asserted where they are)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REL = 1e-5


@pytest.fixture(scope="module")
def im():
    import astroburst_b200 as ab

    ab.default_context()
    return ab.imaging


def sky(rows, cols, seed=0, pad=True):
    rng = np.random.default_rng(seed)
    img = rng.normal(1000.0, 12.0, (rows, cols))
    n = rows * cols // 400
    img.reshape(-1)[rng.integers(0, rows * cols, n)] += rng.uniform(100, 40000, n)
    img = img.astype(np.float32)
    if pad:
        b = max(rows // 40, 1)
        img[:b] = 0.0
        img[:, :b] = 0.0
        if rows > 8 and cols > 8:
            img[5, 7] = np.nan
            img[6, 7] = np.inf
            img[7, 7] = -3.0
    return img


def close(a, b, rel=REL):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    ok = (a == b) | (np.abs(a - b) <= rel * np.abs(b)) | (np.isnan(a) & np.isnan(b))
    return bool(ok.all())


def stats_equal(a, b, mean_rel=1e-12):
    assert a.valid_count == b.valid_count
    assert a.min == b.min and a.max == b.max
    assert a.median == b.median, (a.median, b.median)
    assert a.mad == b.mad and a.sigma == b.sigma
    assert abs(a.mean - b.mean) <= mean_rel * abs(b.mean)  # f64 sum order


# ------------------------------------------------------------------ stats
@pytest.mark.parametrize("shape", [(4, 4), (100, 100), (257, 513), (1600, 1600)])
def test_stats_exact_path(im, orc, shape):
    d = sky(*shape, seed=shape[0])
    stats_equal(im.compute_image_stats(d), orc.compute_image_stats(d))
    d2 = d.copy().reshape(-1)[:-1]  # odd / even valid counts both occur
    stats_equal(im.compute_image_stats(d2), orc.compute_image_stats(d2))


def test_stats_reference_cases(im, orc):
    raw = np.zeros(10000, np.float32)
    raw[3750:6250] = np.arange(1, 2501, dtype=np.float32) * np.float32(0.001)
    st = im.compute_image_stats(raw)   # stf.rs:185-201
    assert st.valid_count == 2500 and st.min > 0.0
    stats_equal(st, orc.compute_image_stats(raw))
    empty = im.compute_image_stats(np.zeros((8, 8), np.float32))
    assert empty.valid_count == 0 and empty.min == 0.0 and empty.sigma == 0.0   # ImageStats::default
    assert im.compute_image_stats(np.zeros(0, np.float32)).valid_count == 0


def test_stats_hist_path(im, orc):
    d = sky(2100, 2000, seed=3)           # 4.2 Mpx > 4,000,000 -> histogram algorithm
    a, b = im.compute_image_stats(d), orc.compute_image_stats(d)
    stats_equal(a, b)
    flat = np.random.default_rng(5).uniform(0.5, 2.0, 4_100_000).astype(np.float32)
    stats_equal(im.compute_image_stats(flat), orc.compute_image_stats(flat))
    stats_equal(im.compute_image_stats_hist(d[:300]), orc.compute_image_stats_hist(d[:300]))
    kr = im.compute_image_stats_with_known_range(d, 10.0, 50000.0)
    stats_equal(kr, orc.compute_image_stats_with_known_range(d, 10.0, 50000.0))
    stats_equal(im.compute_image_stats_with_known_range(d, 5.0, 5.0), b)      # bad range -> full path
    small = d[:100]
    stats_equal(im.compute_image_stats_with_known_range(small, 0.0, 1.0), orc.compute_image_stats(small))


def test_histogram_bins_bit_exact(im, orc):
    d = sky(900, 1100, seed=9)
    st = orc.compute_image_stats(d)
    bins, edges = im.build_histogram(d, 65536, st.min, st.max)
    obins, oedges = orc.build_histogram(d, 65536, st.min, st.max)
    assert np.array_equal(bins, obins) and np.array_equal(edges, oedges)
    assert int(bins.sum()) == st.valid_count
    assert np.array_equal(im.downsample_histogram(bins, 512), orc.downsample_histogram(obins, 512))
    z, ze = im.build_histogram(d, 65536, 5.0, 5.0)           # range < 1e-10 (stats.rs:379-387)
    assert not z.any() and (ze == 5.0).all()
    flat = np.random.default_rng(2).uniform(0.0, 1.0, 300000).astype(np.float32)
    b2, _ = im.build_histogram(flat, 65536, 0.1, 0.9)        # out-of-range pixels clamp / floor to 0
    o2, _ = orc.build_histogram(flat, 65536, 0.1, 0.9)
    assert np.array_equal(b2, o2)


def test_auto_stf_and_wb(im, orc):
    for seed in range(4):
        d = sky(300, 300, seed=seed)
        a, b = im.compute_image_stats(d), orc.compute_image_stats(d)
        for k in (-2.8, -1.5, -4.0):
            p, q = im.auto_stf(a, 0.25, k), orc.auto_stf(b, 0.25, k)
            assert (p.shadow, p.midtone, p.highlight) == (q.shadow, q.midtone, q.highlight)
    r, g, b = (orc.compute_image_stats(sky(200, 200, seed=s) * np.float32(f)) for s, f in ((1, 1.0), (2, 0.7), (3, 1.4)))
    r2, g2, b2 = (im.ImageStats(x.min, x.max, x.median, x.mad, x.sigma, x.mean, x.valid_count) for x in (r, g, b))
    assert im.select_wb_reference(r2, g2, b2) == orc.select_wb_reference(r, g, b)
    c, oc = im.linked_stats(r2, g2, b2), orc.linked_stats(r, g, b)
    assert (c.min, c.max, c.mean, c.median, c.sigma, c.mad) == (oc.min, oc.max, oc.mean, oc.median, oc.sigma, oc.mad)


# ------------------------------------------------------------------ STF / stretch
def test_apply_stf_u8_bit_exact(im, orc):
    d = sky(700, 900, seed=4)
    st = orc.compute_image_stats(d)
    ist = im.ImageStats(st.min, st.max, st.median, st.mad, st.sigma, st.mean, st.valid_count)
    for p in (orc.auto_stf(st), orc.StfParams(0.0, 0.5, 1.0), orc.StfParams(0.1, 0.2, 0.9)):
        ip = im.StfParams(p.shadow, p.midtone, p.highlight)
        assert np.array_equal(im.apply_stf(d, ip, ist), orc.apply_stf(d, p, st))
        a, b = im.apply_stf_f32(d, ip, ist), orc.apply_stf_f32(d, p, st)
        assert close(a, b) and (a != b).mean() < 1e-3
        assert close(im.apply_stf_f32(d, ip, ist, rgb_variant=True), orc.apply_stf_f32(d, p, st, rgb_variant=True))
    data = (np.arange(1, 17, dtype=np.float32) * 100.0).reshape(4, 4)      # stf.rs:216-229
    st = im.compute_image_stats(data)
    buf = im.apply_stf(data, im.StfParams(0.0, 0.5, 1.0), st).reshape(-1)
    assert buf[0] == 0 and buf[15] == 255


def test_arcsinh(im, orc):
    d = sky(500, 600, seed=6)
    for factor, gamma in ((10.0, 1.0), (50.0, 1.0), (500.0, 0.5), (1.0, 2.2)):
        a = im.arcsinh_stretch_with_stats(d, 900.0, 30000.0, factor, gamma)
        b = orc.arcsinh_stretch_with_stats(d, 900.0, 30000.0, factor, gamma)
        # north_star tolerance 1e-5 relative; measured on B200: max rel 4.8e-7, max abs 2.4e-7 (libm vs CUDA
        # log1pf / hypotf / powf, <= 2 ulp of [0, 1])
        assert close(a, b, 1e-5) and np.abs(a - b).max() < 1e-6
    assert np.array_equal(im.arcsinh_stretch_with_stats(d, 0.0, 1.0, 0.0), d, equal_nan=True)   # factor 0 -> clone
    assert not im.arcsinh_stretch_with_stats(d, 5.0, 5.0, 10.0).any()                            # range < 1e-10 -> zeros
    r = im.arcsinh_stretch_with_stats(np.array([[np.nan, -0.5, 0.5]], np.float32), 0.0, 1.0, 10.0)
    assert r[0, 0] == 0.0 and r[0, 1] == 0.0 and r[0, 2] > 0.0
    ch = (np.add.outer(np.arange(10), np.arange(10)) / 20.0).astype(np.float32)      # stretch.rs:146-153
    single = im.arcsinh_stretch(ch, 20.0)
    for o in im.arcsinh_stretch_rgb(ch, ch, ch, 20.0):
        assert np.array_equal(single, o)


# ------------------------------------------------------------------ blend / SCNR / scale / tone
def test_blend_channels_bit_exact(im, orc):
    chans = [sky(300, 500, seed=s, pad=False) for s in range(3)]
    sho = [(0, 1.0, 0.0, 0.0), (1, 0.0, 1.0, 0.0), (2, 0.0, 0.0, 1.0)]
    hubble = [(0, 0.8, 0.1, 0.0), (1, 0.2, 0.7, 0.15), (2, 0.0, 0.2, 0.85), (7, 1.0, 1.0, 1.0)]
    for w in (sho, hubble):
        a = im.blend_channels(chans, w, 300, 500)
        b = orc.blend_channels(chans, w, 300, 500)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)          # mul-then-add in f32, no FMA: bit-exact
    chans[1][10, 10] = np.nan                    # no NaN filtering (G5)
    a = im.blend_channels(chans, hubble, 300, 500)
    assert np.isnan(a[0][10, 10]) and np.isnan(a[1][10, 10])
    short = [chans[0], chans[1].reshape(-1)[:1001], chans[2]]   # pixels beyond a channel's length
    a, b = im.blend_channels(short, hubble, 300, 500), orc.blend_channels(short, hubble, 300, 500)
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)


@pytest.mark.parametrize("method", ["average", "maximum"])
@pytest.mark.parametrize("preserve", [False, True])
def test_scnr_bit_exact(im, orc, method, preserve):
    rng = np.random.default_rng(11)
    base = [rng.uniform(0.0, 1.3, (257, 301)).astype(np.float32) for _ in range(3)]
    for amount in (1.0, 0.5, 0.0, 1e-8, 2.0):
        a = [x.copy() for x in base]
        b = [x.copy() for x in base]
        im.apply_scnr_inplace(*a, method, amount, preserve)
        orc.apply_scnr_inplace(*b, method, amount, preserve)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    a = [base[0].copy(), base[1][:200].copy(), base[2].copy()]       # dims differ -> silent no-op
    im.apply_scnr_inplace(*a)
    assert np.array_equal(a[1], base[1][:200])


def test_scale_levels_curves(im, orc):
    d = sky(400, 333, seed=8)
    assert np.array_equal(im.scale(d, 1.37), orc.scale(d, 1.37), equal_nan=True)
    u = np.random.default_rng(3).uniform(-0.2, 1.4, (300, 257)).astype(np.float32)
    u[0, 0] = np.nan
    for lv in ((0.0, 1.0, 1.0), (0.2, 1.0, 1.0), (0.05, 2.0, 0.9), (0.0, 0.5, 1.0), (0.1, 100.0, 0.1)):
        a, b = im.apply_levels(u, *lv), orc.apply_levels(u, *lv)
        assert close(a, b, 2e-6) if lv != (0.0, 1.0, 1.0) else np.array_equal(a, b, equal_nan=True)
    for pts in ([(0.0, 0.0), (1.0, 1.0)], [(0.0, 0.0), (0.25, 0.15), (0.5, 0.5), (0.75, 0.85), (1.0, 1.0)],
                [(0.3, 0.1), (0.7, 0.9)], [], [(0.5, 0.5), (0.5, 0.6), (0.2, 0.9)]):
        lut, olut = im.spline_lut(pts), orc.spline_lut(pts)
        assert np.array_equal(lut, olut)
        assert np.array_equal(im.apply_curve(u, lut), orc.apply_curve(u, olut))


def test_tone_composite_fused_equals_separate_steps(im, orc):
    """apply_tone_composite_cmd's chain (curves.rs:111-151) in one sweep == the oracle's four
    separate operators, per-channel and linked STF, with and without levels / curves / SCNR."""
    chans = [sky(257, 301, seed=20 + k) * np.float32(1.0 + 0.1 * k) for k in range(3)]
    ost = [orc.compute_image_stats(c) for c in chans]
    gst = [im.compute_image_stats(c) for c in chans]
    pts = [(0.0, 0.0), (0.25, 0.15), (0.5, 0.5), (0.75, 0.85), (1.0, 1.0)]
    ident = [(0.0, 0.0), (1.0, 1.0)]
    cases = [
        dict(linked=False, levels=None, curves=None, scnr=None),
        dict(linked=True, levels=None, curves=[pts, ident, pts], scnr=("average", 0.5, False)),
        dict(linked=False, levels=[(0.02, 1.4, 0.95), (0.0, 1.0, 1.0), (0.1, 0.7, 1.0)], curves=None,
             scnr=("maximum", 1.0, True)),
        dict(linked=True, levels=[(0.05, 2.0, 0.9)] * 3, curves=[pts] * 3, scnr=("average", 1e-9, True)),
    ]
    for c in cases:
        if c["linked"]:
            ocomb = orc.linked_stats(*ost)
            ostf, onorm = [orc.auto_stf(ocomb)] * 3, [ocomb] * 3
            gcomb = im.linked_stats(*gst)
            gstf, gnorm = [im.auto_stf(gcomb)] * 3, [gcomb] * 3
        else:
            ostf, onorm = [orc.auto_stf(s_) for s_ in ost], ost
            gstf, gnorm = [im.auto_stf(s_) for s_ in gst], gst
        want = [orc.apply_stf_f32(ch, p, n) for ch, p, n in zip(chans, ostf, onorm)]
        if c["levels"]:
            want = [orc.apply_levels(w, *lv) for w, lv in zip(want, c["levels"])]
        luts = None
        if c["curves"]:
            luts = [im.spline_lut(p) for p in c["curves"]]
            want = [orc.apply_curve(w, orc.spline_lut(p)) for w, p in zip(want, c["curves"])]
        if c["scnr"] and c["scnr"][1] > 1e-7:
            want = [np.ascontiguousarray(w) for w in want]
            orc.apply_scnr_inplace(*want, *c["scnr"])
        got = im.apply_tone_composite(*chans, gstf, gnorm, c["levels"], luts, c["scnr"])
        for x, y in zip(got, want):
            if c["levels"] and c["curves"]:
                # an ulp of f64 pow (libm vs CUDA) can move a pixel across one LUT index
                d = np.abs(x.astype(np.float64) - y)
                off = d > 2e-6 * np.abs(y)
                assert off.mean() < 1e-3 and (not off.any() or d[off].max() < 2e-3)
            elif c["levels"]:
                assert close(x, y, 2e-6)       # f64 pow: within the libm/CUDA ulp difference
            else:
                assert np.array_equal(x, y)


# ------------------------------------------------------------------ bicubic gathers
def test_warp_shift_resample_bit_exact(im, orc):
    img = sky(301, 407, seed=12, pad=False)
    import math
    th = math.radians(1.7)
    for t in ((1, 0, 0, 0, 1, 0), (1, 0, 5.0, 0, 1, 3.0), (1.004 * math.cos(th), -math.sin(th), 12.3, math.sin(th), 0.997 * math.cos(th), -7.9),
              (1, 0, 1000.0, 0, 1, 1000.0), (0.5, 0, 0.25, 0, 0.5, 0.75)):
        for shape in ((301, 407), (200, 500)):
            a, b = im.warp_image(img, t, *shape), orc.warp_image(img, t, *shape)
            assert np.array_equal(a, b), (t, shape, np.abs(a - b).max())
    for dy, dx in ((0.0, 0.0), (1e-13, 0.0), (0.5, -0.5), (3.25, 7.75), (-400.0, 2.0), (0.3, 406.6)):
        assert np.array_equal(im.shift_image_subpixel(img, dy, dx), orc.shift_image_subpixel(img, dy, dx))
    for tr, tc in ((301, 407), (150, 203), (602, 814), (77, 1000), (1, 1)):
        assert np.array_equal(im.resample_image(img, tr, tc), orc.resample_image(img, tr, tc))
    with pytest.raises(Exception):
        im.resample_image(img, 0, 10)
    ramp = np.arange(2500, dtype=np.float32).reshape(50, 50)                       # affine.rs:789-800
    w = im.warp_image(ramp, (1, 0, 0, 0, 1, 0), 50, 50)
    assert np.abs(w[2:48, 2:48] - ramp[2:48, 2:48]).max() < 0.5


def test_bicubic_weight_joints_and_odd_inputs(im, orc):
    """The statically chosen Catmull-Rom pieces against the reference's comparisons where they could
    differ: fractions that round onto 0 / 1 (taps exactly on the joints |t| = 1, 2), integer
    coordinates, non-finite samples and transforms, planes smaller than the 4 x 4 footprint."""
    img = sky(64, 80, seed=3, pad=False)
    eps = np.finfo(np.float64).eps
    for dy, dx in ((-1e-20, -1e-20), (-eps / 4, eps / 4), (1 - eps / 2, 1 - eps / 2), (-0.5, -0.5), (2.0, 3.0),
                   (1e-300, -1e-300), (0.5 - eps, 63.5), (-0.49999999999999994, 79.49999999999999)):
        assert np.array_equal(im.shift_image_subpixel(img, dy, dx), orc.shift_image_subpixel(img, dy, dx)), (dy, dx)
    for dy, dx in ((float("nan"), 0.25), (0.25, float("nan")), (float("inf"), 0.0), (0.0, float("-inf"))):
        a, b = im.shift_image_subpixel(img, dy, dx), orc.shift_image_subpixel(img, dy, dx)   # NaN shifts pass the bounds rule
        assert np.array_equal(a, b, equal_nan=True) and np.array_equal(np.signbit(a), np.signbit(b)), (dy, dx)
    for t in ((1, 0, 1e-17, 0, 1, 1e-17), (1, 0, 1 - eps / 2, 0, 1, 2 - eps), (1, 1e-18, 0, 1e-18, 1, 0),
              (1, 0, float("nan"), 0, 1, 0), (float("inf"), 0, 0, 0, 1, 0), (0, 0, 5.0, 0, 0, 7.0), (-1, 0, 70.0, 0, -1, 60.0)):
        a, b = im.warp_image(img, t, 64, 80), orc.warp_image(img, t, 64, 80)
        assert np.array_equal(a, b, equal_nan=True), t
    bad = img.copy()
    bad[10, 10], bad[20, 30], bad[40, 5] = np.nan, np.inf, -np.inf
    for t in ((1, 0, 0.25, 0, 1, 0.5), (1, 0, 0, 0, 1, 0)):
        a, b = im.warp_image(bad, t, 64, 80), orc.warp_image(bad, t, 64, 80)
        assert np.array_equal(a, b, equal_nan=True) and np.array_equal(np.signbit(a), np.signbit(b))
    assert np.array_equal(im.shift_image_subpixel(bad, 0.3, -0.7), orc.shift_image_subpixel(bad, 0.3, -0.7), equal_nan=True)
    for shape in ((1, 1), (2, 3), (3, 2), (4, 4), (1, 9)):
        tiny = sky(*shape, seed=5, pad=False)
        assert np.array_equal(im.shift_image_subpixel(tiny, 0.25, -0.25), orc.shift_image_subpixel(tiny, 0.25, -0.25))
        assert np.array_equal(im.resample_image(tiny, 7, 5), orc.resample_image(tiny, 7, 5))
        assert np.array_equal(im.warp_image(tiny, (1, 0, 0.5, 0, 1, 0.5), 6, 6), orc.warp_image(tiny, (1, 0, 0.5, 0, 1, 0.5), 6, 6))
    zeros = np.zeros((32, 32), np.float32)
    neg = -sky(32, 32, seed=6, pad=False)
    for plane in (zeros, neg):   # signed zeros of the result
        a, b = im.shift_image_subpixel(plane, 0.5, 0.5), orc.shift_image_subpixel(plane, 0.5, 0.5)
        assert np.array_equal(a, b) and np.array_equal(np.signbit(a), np.signbit(b))
