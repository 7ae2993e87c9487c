"""GPU parity tests for tile background, star detection and the affine aligner vs the CPU oracle.

Bars: background median / sigma, star count, order, npix bit-exact; star parameters bit-exact
(the sparse flood fill keeps the reference's summation order); the aligner's transform equals the
oracle's for the same emulated rayon thread count (the reference itself is nondeterministic here:
HashMap vote-tie order and thread-count dependent RANSAC streams), and is also checked
geometrically against the known ground-truth transform."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def al():
    import astroburst_b200 as ab

    ab.default_context()
    return ab.alignment


def star_field(rows, cols, n_stars, seed, transform=None, noise_seed=0, sky=100.0):
    """Gaussian stars (sigma 1.6 px) on a noisy sky; `transform` maps reference -> this frame."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(20, cols - 20, n_stars)
    y = rng.uniform(20, rows - 20, n_stars)
    flux = 200.0 * (30000.0 / 200.0) ** (rng.uniform(0, 1, n_stars) ** 2)
    if transform is not None:
        a, b, tx, c, d, ty = transform
        x, y = a * x + b * y + tx, c * x + d * y + ty
    img = np.random.default_rng(1000 + noise_seed).normal(sky, 3.0, (rows, cols))
    r = 8
    for sx, sy, f in zip(x, y, flux):
        cx, cy = int(round(sx)), int(round(sy))
        if cx - r < 0 or cy - r < 0 or cx + r >= cols or cy + r >= rows:
            continue
        yy, xx = np.mgrid[cy - r:cy + r + 1, cx - r:cx + r + 1]
        img[cy - r:cy + r + 1, cx - r:cx + r + 1] += f / (2 * np.pi * 1.6 ** 2) * np.exp(
            -((xx - sx) ** 2 + (yy - sy) ** 2) / (2 * 1.6 ** 2))
    return img.astype(np.float32)


def reference_test_image(rows, cols):   # star_detection.rs:264-287
    r, c = np.mgrid[0:rows, 0:cols]
    img = (100.0 + ((r * 7 + c * 13) % 17).astype(np.float32) * np.float32(0.5)).astype(np.float32)
    for sy, sx, peak in [(50, 50, 5000.0), (100, 200, 3000.0), (200, 150, 8000.0)]:
        for dy in range(-5, 6):
            for dx in range(-5, 6):
                rr, cc = sy + dy, sx + dx
                if 0 <= rr < rows and 0 <= cc < cols:
                    img[rr, cc] = np.float32(img[rr, cc] + np.float32(peak * math.exp(-(dx * dx + dy * dy) / 8.0)))
    return img


def same_detection(a, b):
    assert a.background_median == b.background_median and a.background_sigma == b.background_sigma
    assert len(a.stars) == len(b.stars)
    for s, t in zip(a.stars, b.stars):
        assert s == t, (s, t)


def test_reference_unit_cases(al, orc):
    img = reference_test_image(300, 300)
    res = al.detect_stars(img, 5.0)
    assert len(res.stars) >= 3 and res.background_sigma > 0.0
    assert res.stars[0]["flux"] >= res.stars[1]["flux"]
    assert abs(res.stars[0]["x"] - 150.0) < 2.0 and abs(res.stars[0]["y"] - 200.0) < 2.0
    same_detection(res, orc.detect_stars(img, 5.0))
    assert al.detect_stars(np.full((100, 100), 50.0, np.float32), 5.0).stars == []
    med, sig = al.estimate_background(np.full((200, 200), 100.0, np.float32), 64)
    assert abs(med - 100.0) < 1.0 and sig < 1.0
    tiny = al.detect_stars(np.ones((2, 50), np.float32), 5.0)      # rows < 3 (star_detection.rs:89-98)
    assert tiny.stars == [] and tiny.background_sigma == 1.0


@pytest.mark.parametrize("shape,tile", [((200, 200), 64), ((333, 517), 40), ((90, 1000), 16), ((1024, 1024), 128)])
def test_background_matches_oracle(al, orc, shape, tile):
    img = star_field(*shape, n_stars=60, seed=shape[0])
    img[:7, :] = 0.0
    img[11, 13] = np.nan
    assert al.estimate_background(img, tile) == orc.estimate_background(img, tile)


@pytest.mark.parametrize("kind", ["integer", "integer/65535", "8 levels", "two values", "constant + outliers",
                                  "mostly invalid", "ramp"])
def test_background_sampled_selection_on_tie_heavy_and_odd_tiles(al, orc, kind):
    """The sampled-bracket selection of the 256^2 on-chip tiles where the bracket degenerates:
    massive ties (integer ADU planes: a handful of distinct keys around the median), two-valued and
    constant tiles, tiles with few valid pixels (thin samples), monotone ramps -- and the radix-select
    fallback they force.  Medians / MADs / clip rounds must equal the oracle's bit for bit."""
    rng = np.random.default_rng(11)
    n = 1024                                  # 16 tiles of 256 x 256 (the detector's tile size here)
    base = 1000.0 + rng.normal(0.0, 12.0, (n, n))
    base[rng.random((n, n)) < 2e-3] += 20000.0
    if kind == "integer":
        img = np.rint(base)
    elif kind == "integer/65535":
        img = np.rint(base) / 65535.0
    elif kind == "8 levels":
        img = np.rint(base / 8.0) * 8.0
    elif kind == "two values":
        img = np.where(rng.random((n, n)) < 0.5, 1000.0, 1001.0)
    elif kind == "constant + outliers":
        img = np.full((n, n), 777.0)
        img[rng.random((n, n)) < 1e-3] = 5000.0
    elif kind == "mostly invalid":
        img = base.copy()
        img[rng.random((n, n)) < 0.98] = 0.0          # ~1,300 valid pixels per tile
        img[:256, :256] = 0.0                          # one empty tile
        img[:256, 256:512][rng.random((256, 256)) < 0.9995] = 0.0   # a tile with ~30 valid pixels
        img[256:512, :256] = np.nan
        img[256:263, :1] = base[256:263, :1]           # 7 valid pixels: below the n >= 8 gate
    else:
        img = np.add.outer(np.arange(n), np.arange(n)).astype(np.float64) * 0.01 + 1.0
    img = img.astype(np.float32)
    assert al.estimate_background(img, 256) == orc.estimate_background(img, 256)
    same_detection(al.detect_stars(img, 4.0), orc.detect_stars(img, 4.0))


@pytest.mark.parametrize("shape,n,sigma", [((512, 512), 80, 5.0), ((700, 900), 300, 3.5), ((1500, 1100), 500, 4.0)])
def test_detect_matches_oracle(al, orc, shape, n, sigma):
    img = star_field(*shape, n_stars=n, seed=n)
    img[:, -9:] = 0.0
    same_detection(al.detect_stars(img, sigma), orc.detect_stars(img, sigma))
    part = al.detect_stars(img, sigma, max_stars=5)
    assert len(part.stars) == 5


def test_normalize_matches_oracle(al, orc):
    img = star_field(400, 600, 100, seed=3)
    img[5, 5] = np.nan
    assert np.array_equal(al.normalize_for_detection(img), orc.normalize_for_detection(img), equal_nan=True)
    small = img[:5, :10]                                            # < 100 samples -> unchanged
    assert np.array_equal(al.normalize_for_detection(small), small)


def test_match_and_ransac_match_oracle(al, orc):
    rng = np.random.default_rng(5)
    ref = rng.uniform(50, 950, (80, 2))
    th = math.radians(1.2)
    A = np.array([[1.003 * math.cos(th), -math.sin(th)], [math.sin(th), 0.998 * math.cos(th)]])
    tgt = ref @ A.T + np.array([12.5, -8.25]) + rng.normal(0, 0.05, ref.shape)
    tgt = tgt[rng.permutation(80)][:70]
    m, om = al.match_stars(ref, tgt), orc.match_stars(ref, tgt)
    assert np.array_equal(m, om) and len(m) >= 20
    for method in ("affine", "rigid"):
        for nt in (1, 8, 16):
            r, o = al.ransac_affine(m, method, nt), orc.ransac(om, method, nt)
            assert (r is None) == (o is None)
            g = al.ransac_affine(m, method, nt, on_gpu=True)   # hypotheses on the device: same bits
            assert (g is None) == (o is None)
            if g is not None:
                assert tuple(getattr(g.transform, k) for k in "a b tx c d ty".split()) == o.transform.as_tuple()
                assert (g.inliers, g.matched_stars, g.residual_px) == (o.inliers, o.matched_stars, o.residual_px)
            if r is not None:
                assert tuple(getattr(r.transform, k) for k in "a b tx c d ty".split()) == o.transform.as_tuple()
                assert (r.inliers, r.matched_stars, r.residual_px) == (o.inliers, o.matched_stars, o.residual_px)
    stars = [(10.0, 10.0), (50.0, 10.0), (30.0, 40.0), (80.0, 20.0), (60.0, 70.0)]     # affine.rs:808-822
    assert len(al.match_stars(stars, stars)) >= 4


@pytest.mark.parametrize("kind", ["uniform", "elongated", "few", "clustered"])
def test_bucketed_triangle_votes_equal_all_pairs(al, orc, kind, monkeypatch):
    """The bucket-sorted vote kernel against the all-pairs kernel (AB_TRI_BRUTE) and the oracle: the
    same matches, also when side ratios run past the last bucket (elongated fields: ratios > 82)."""
    rng = np.random.default_rng({"uniform": 1, "elongated": 2, "few": 3, "clustered": 4}[kind])
    if kind == "uniform":
        ref = rng.uniform(20, 4000, (120, 2))
    elif kind == "elongated":   # a thin strip: mid / shortest ratios of several hundred
        ref = np.stack([rng.uniform(0, 8000, 70), rng.uniform(0, 40, 70)], axis=1)
    elif kind == "few":
        ref = rng.uniform(20, 600, (7, 2))
    else:
        ref = np.concatenate([rng.normal(500, 30, (40, 2)), rng.uniform(0, 3000, (30, 2))])
    tgt = ref * 1.0005 + np.array([3.25, -7.5]) + rng.normal(0, 0.08, ref.shape)
    tgt = tgt[rng.permutation(len(tgt))]
    fast = al.match_stars(ref, tgt)
    monkeypatch.setenv("AB_TRI_BRUTE", "1")
    brute = al.match_stars(ref, tgt)
    monkeypatch.delenv("AB_TRI_BRUTE")
    assert np.array_equal(fast, brute)
    assert np.array_equal(fast, orc.match_stars(ref, tgt))
    assert len(fast) >= 3


@pytest.mark.parametrize("seed,n,outliers,noise", [(1, 60, 0, 0.05), (2, 60, 25, 0.3), (3, 9, 3, 0.1), (4, 6, 0, 0.0),
                                                   (5, 120, 80, 1.0), (6, 40, 38, 0.2), (7, 3, 0, 0.0)])
def test_device_ransac_matches_oracle(al, orc, seed, n, outliers, noise):
    """ab_ransac_affine_gpu against the oracle on clean, contaminated, tiny and hopeless match lists."""
    rng = np.random.default_rng(seed)
    ref = rng.uniform(0, 4000, (n, 2))
    A = np.array([[0.999, -0.012], [0.011, 1.002]])
    tgt = ref @ A.T + np.array([-20.5, 14.25]) + rng.normal(0, noise, ref.shape)
    tgt[:outliers] = rng.uniform(0, 4000, (outliers, 2))
    m = np.concatenate([ref, tgt], axis=1)
    for nt in (1, 8, 7):
        g, o = al.ransac_affine(m, "affine", nt, on_gpu=True), orc.ransac(m, "affine", nt)
        assert (g is None) == (o is None)
        if g is not None:
            assert tuple(getattr(g.transform, k) for k in "a b tx c d ty".split()) == o.transform.as_tuple()
            assert (g.inliers, g.matched_stars, g.residual_px) == (o.inliers, o.matched_stars, o.residual_px)


def test_align_recovers_known_transform(al, orc):
    rows, cols = 1024, 1024
    th = math.radians(0.8)
    t_true = (1.002 * math.cos(th), -math.sin(th), 9.5, math.sin(th), 1.002 * math.cos(th), -6.25)
    ref = star_field(rows, cols, 250, seed=11)
    tgt = star_field(rows, cols, 250, seed=11, transform=t_true, noise_seed=1)
    for nt in (8, 16):
        r, o = al.align_channel_affine(ref, tgt, nt), orc.align_channel_affine(ref, tgt, nt)
        got = tuple(getattr(r.transform, k) for k in "a b tx c d ty".split())
        assert got == o.transform.as_tuple() and r.method == o.method and r.inliers == o.inliers
        assert r.method in (0, 1) and r.inliers >= 4
        # geometric check: corner displacement vs ground truth below 1 px
        for x, y in ((0, 0), (cols, 0), (0, rows), (cols, rows)):
            gx, gy = got[0] * x + got[1] * y + got[2], got[3] * x + got[4] * y + got[5]
            ex, ey = t_true[0] * x + t_true[1] * y + t_true[2], t_true[3] * x + t_true[4] * y + t_true[5]
            assert math.hypot(gx - ex, gy - ey) < 1.0
    flat = np.full((256, 256), 10.0, np.float32)                   # no stars -> identity fallback
    r = al.align_channel_affine(flat, flat)
    assert r.method == 3 and (r.transform.a, r.transform.tx) == (1.0, 0.0)


def test_align_falls_back_to_phase_correlation(al, orc):
    """No detectable stars but a correlatable texture: affine.rs:243-270 -> method 2 (pure
    translation from phase_correlate), identical decision and ~1e-6 px agreement with the oracle."""
    rng = np.random.default_rng(5)
    base = np.cumsum(np.cumsum(rng.normal(0, 1, (300, 340)), axis=0), axis=1).astype(np.float32)
    base = (base - base.min()) / (base.max() - base.min()) * 100.0 + 50.0   # smooth, star-free
    ref, tgt = base[10:266, 12:268].copy(), base[14:270, 7:263].copy()
    r, o = al.align_channel_affine(ref, tgt, 8), orc.align_channel_affine(ref, tgt, 8)
    assert r.method == o.method
    got, want = (r.transform.a, r.transform.b, r.transform.tx, r.transform.c, r.transform.d, r.transform.ty), \
        o.transform.as_tuple()
    assert all(abs(g - w) < 1e-6 for g, w in zip(got, want)), (got, want)
    if r.method == 2:
        assert (r.transform.a, r.transform.b, r.transform.c, r.transform.d) == (1.0, 0.0, 0.0, 1.0)
        assert r.matched_stars == 0 and r.inliers == 0


def test_detect_components_on_device_are_order_exact(al, orc):
    """Connected components, visiting order and f64 moments now run on the device (union-find CCL +
    a replay of the reference's BFS per component): star lists must equal the oracle's -- count,
    order, npix and every f64 parameter -- on shapes where the visiting order matters."""
    img = star_field(600, 700, 150, seed=21)
    img[100:140, 200:260] += 500.0            # a 2400-px blob (kept: <= 5000 px)
    img[300:380, 400:480] += 400.0            # a 6400-px blob (flood-filled, then dropped: > 5000 px)
    img[200:230, 100] += 300.0                # U shape: two arms joined at the bottom -> the seed is
    img[200:230, 110] += 300.0                # not the only raster-first pixel of its component
    img[229, 100:111] += 300.0
    img[0:4, 300:305] += 350.0                # touches the top border (seeded from row 1)
    img[500:520, 696:700] += 350.0            # touches the right border
    img[-1, 10:40] += 350.0                   # border-only pixels: never seeded
    img[450, 50:53] += 200.0                  # exactly 3 px
    img[460, 60:62] += 200.0                  # 2 px: dropped
    img[470:475:2, 70:75:2] += 200.0          # diagonal-only contacts are NOT connected at distance 2
    img[480, 80] += 200.0                     # 8-connectivity: a diagonal chain is one component
    img[481, 81] += 200.0
    img[482, 82] += 200.0
    img[483, 81] += 200.0
    want = orc.detect_stars(img, 3.5)
    assert len(want.stars) > 50
    same_detection(al.detect_stars(img, 3.5), want)
    for sigma in (2.0, 5.0, 12.0):
        same_detection(al.detect_stars(img, sigma), orc.detect_stars(img, sigma))
    # dense noise just above the threshold: thousands of tiny irregular components
    rng = np.random.default_rng(3)
    noisy = rng.normal(100.0, 5.0, (512, 640)).astype(np.float32)
    same_detection(al.detect_stars(noisy, 1.0), orc.detect_stars(noisy, 1.0))
    same_detection(al.detect_stars(noisy, 0.3), orc.detect_stars(noisy, 0.3))   # one percolating blob
    # nothing above the threshold / everything above it
    flat = np.full((64, 80), 7.0, np.float32)
    same_detection(al.detect_stars(flat, 3.5), orc.detect_stars(flat, 3.5))
    step = flat.copy()
    step[:, 40:] = 500.0
    same_detection(al.detect_stars(step, 3.5), orc.detect_stars(step, 3.5))


@pytest.mark.parametrize("shape,tile", [((700, 900), 300), ((1100, 600), 512)])
def test_background_tiles_beyond_the_on_chip_capacity(al, orc, shape, tile):
    """Caller-chosen tiles above 65,536 px run the global-memory pixel source of the tile kernel
    (the detector's own tiles, <= 256 x 256, always run from the on-chip copy)."""
    img = star_field(*shape, n_stars=80, seed=shape[1])
    img[:9, :] = 0.0
    img[13, 17] = np.nan
    assert al.estimate_background(img, tile) == orc.estimate_background(img, tile)


# ---- subframe scoring (core/analysis/subframe.rs:52-158)
def test_analyze_subframe_matches_oracle(al, orc):
    from astroburst_b200 import alignment

    imgs = [star_field(400, 500, 120, seed=4), star_field(300, 300, 40, seed=9),
            np.full((64, 64), 10.0, np.float32), reference_test_image(300, 300)]
    got = []
    for img in imgs:
        g, w = alignment.analyze_subframe(img), orc.analyze_subframe(img)
        assert g.as_tuple() == w.as_tuple()
        got.append(g)
    cfg_g = alignment.SubframeWeightConfig(max_fwhm=2.0, min_stars=500, snr_weight=0.0)
    cfg_o = orc.SubframeConfig(max_fwhm=2.0, min_stars=500, snr_weight=0.0)
    assert alignment.analyze_subframe(imgs[0], cfg_g).as_tuple() == orc.analyze_subframe(imgs[0], cfg_o).as_tuple()
    want = orc.normalize_weights([g.weight for g in got])
    alignment.normalize_weights(got)
    assert [g.weight for g in got] == want and max(want) == 1.0
