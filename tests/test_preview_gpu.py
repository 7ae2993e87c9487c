"""GPU parity tests for the preview/export tail, the LRGB merge, the star mask and the masked
stretch: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

Bars: u8 / u16 preview bytes bit-exact (integer work); LRGB, luminance, star mask and masked
stretch f32 pixels bit-exact (same f32 expression order, no FMA); masked medians bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pv():
    import astroburst_b200 as ab

    ab.default_context()
    return ab.preview


@pytest.fixture(scope="module")
def im():
    import astroburst_b200 as ab

    return ab.imaging


def planes01(rows, cols, seed):
    """Three planes mostly in [0, 1] with out-of-range values, NaN and inf sprinkled in."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(3):
        p = rng.uniform(-0.1, 1.1, (rows, cols)).astype(np.float32)
        if rows * cols > 16:
            p.reshape(-1)[rng.integers(0, rows * cols, 4)] = [np.nan, np.inf, -np.inf, 2.5]
        out.append(p)
    return out


def star_field(rows, cols, n_stars, seed, sky=100.0, sat=False):
    rng = np.random.default_rng(seed)
    x = rng.uniform(5, cols - 5, n_stars)
    y = rng.uniform(5, rows - 5, n_stars)
    flux = 300.0 * (60000.0 / 300.0) ** (rng.uniform(0, 1, n_stars) ** 2)
    img = rng.normal(sky, 3.0, (rows, cols))
    r = 8
    for sx, sy, f in zip(x, y, flux):
        cx, cy = int(round(sx)), int(round(sy))
        x0, x1, y0, y1 = max(cx - r, 0), min(cx + r + 1, cols), max(cy - r, 0), min(cy + r + 1, rows)
        yy, xx = np.mgrid[y0:y1, x0:x1]
        img[y0:y1, x0:x1] += f / (2 * np.pi * 1.6 ** 2) * np.exp(
            -((xx - sx) ** 2 + (yy - sy) ** 2) / (2 * 1.6 ** 2))
    img = (img / 4000.0).astype(np.float32)  # sky 0.025, bright stars above the 0.85 ceiling
    if sat:
        img[3, 4] = np.nan
        img[:2] = 0.0
    return img


# ------------------------------------------------------------------ RGB8 / RGB16
@pytest.mark.parametrize("shape,max_dim", [((1, 1), 0), ((7, 5), 0), ((33, 129), 0), ((257, 513), 0),
                                           ((257, 513), 2048), ((257, 513), 100), ((300, 90), 64),
                                           ((64, 64), 64), ((65, 64), 64), ((1600, 1600), 512)])
def test_render_rgb_u8_bit_exact(pv, orc, shape, max_dim):
    r, g, b = planes01(*shape, seed=shape[0] + max_dim)
    got = pv.render_rgb(r, g, b, max_dim)
    want = orc.render_rgb_u8(r, g, b, max_dim)
    assert got.shape == want.shape
    assert pv.preview_size(shape[0], shape[1], max_dim) == orc.preview_size(shape[0], shape[1], max_dim)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (64, 65), (257, 513)])
def test_render_rgb_16bit_bit_exact(pv, orc, shape):
    r, g, b = planes01(*shape, seed=7 + shape[1])
    assert np.array_equal(pv.render_rgb_16bit(r, g, b), orc.render_rgb_u16be(r, g, b))


@pytest.mark.parametrize("shape,max_dim", [((5, 3), 0), ((257, 513), 0), ((257, 513), 128),
                                           ((1600, 1600), 2048), ((1601, 900), 800)])
def test_render_rgb_with_stf_bit_exact(pv, im, orc, shape, max_dim):
    rng = np.random.default_rng(shape[0])
    chans = []
    for k in range(3):
        p = rng.normal(1000.0 + 150 * k, 20.0, shape)
        p.reshape(-1)[rng.integers(0, p.size, max(p.size // 300, 1))] += rng.uniform(100, 30000)
        p = p.astype(np.float32)
        if p.size > 8:
            p.reshape(-1)[:3] = [0.0, np.nan, -5.0]
        chans.append(p)
    ost = [orc.compute_image_stats(c) for c in chans]
    ostf = [orc.auto_stf(s) for s in ost]
    gst = [im.compute_image_stats(c) for c in chans]
    gstf = [im.auto_stf(s) for s in gst]
    got = pv.render_rgb_preview_with_stf(*chans, gstf, gst, max_dim)
    want = orc.render_rgb_stf_u8(*chans, ostf, ost, max_dim)
    assert np.array_equal(got, want)


# ------------------------------------------------------------------ LRGB / luminance
@pytest.mark.parametrize("shape", [(1, 1), (3, 7), (10, 10), (255, 513)])
@pytest.mark.parametrize("lw,cw", [(1.0, 1.0), (0.7, 0.4), (0.0, 1.0)])
def test_apply_lrgb_bit_exact(pv, orc, shape, lw, cw):
    r, g, b = planes01(*shape, seed=shape[1])
    l = np.random.default_rng(3).uniform(0, 1, shape).astype(np.float32)
    if r.size > 4:
        r.reshape(-1)[0] = g.reshape(-1)[0] = b.reshape(-1)[0] = 0.0  # lum_old < 1e-10 branch
    a = [x.copy() for x in (r, g, b)]
    w = [x.copy() for x in (r, g, b)]
    pv.apply_lrgb(l, *a, lw, cw)
    orc.apply_lrgb(l, *w, lw, cw)
    for x, y in zip(a, w):
        assert np.array_equal(x, y, equal_nan=True)


def test_apply_lrgb_dimension_mismatch(pv):
    l = np.full((10, 10), 0.5, np.float32)
    r = np.full((10, 20), 0.5, np.float32)
    with pytest.raises(ValueError):
        pv.apply_lrgb(l, r, r.copy(), r.copy())  # lrgb.rs test_lrgb_dimension_mismatch


@pytest.mark.parametrize("finite_only", [False, True])
def test_luminance_bit_exact(pv, orc, finite_only):
    r, g, b = planes01(129, 67, seed=11)
    got = pv.synthesize_luminance(r, g, b, finite_only)
    want = orc.synthesize_luminance(r, g, b, finite_only)
    assert np.array_equal(got, want, equal_nan=True)


# ------------------------------------------------------------------ star mask
@pytest.mark.parametrize("protect", [False, True])
def test_star_mask_from_detection_bit_exact(pv, orc, protect):
    img = star_field(300, 400, 60, seed=2)
    det = orc.detect_stars(img, 5.0)
    assert len(det.stars) > 20
    stars = det.stars + [dict(x=-50.0, y=10.0, fwhm=3.0), dict(x=399.5, y=299.9, fwhm=4.0),
                         dict(x=10.0, y=10.0, fwhm=0.5), dict(x=float("nan"), y=5.0, fwhm=3.0)]
    img01 = img
    cfg = pv.StarMaskConfig(luminance_protect=protect, luminance_ceiling=0.2)
    got = pv.generate_star_mask_from_detection(img01, stars, cfg)
    want, n, cov = orc.star_mask_from_stars(img01, stars, luminance_protect=protect,
                                            luminance_ceiling=0.2)
    assert got.stars_masked == n
    assert got.coverage_fraction == cov
    assert np.array_equal(got.mask, want)
    assert got.mask.max() == 1.0 and 0.0 < cov < 1.0


def test_generate_star_mask_matches_oracle(pv, orc):
    img = star_field(257, 321, 40, seed=5)
    got = pv.generate_star_mask(img)
    want, n, cov = orc.generate_star_mask(img)
    assert (got.stars_masked, got.coverage_fraction) == (n, cov)
    assert np.array_equal(got.mask, want)


def test_star_mask_no_stars(pv, orc):
    img = np.full((40, 50), 0.9, np.float32)
    got = pv.generate_star_mask_from_detection(img, [], pv.StarMaskConfig(luminance_protect=True))
    want, n, cov = orc.star_mask_from_stars(img, [], luminance_protect=True)
    assert got.stars_masked == 0 and np.array_equal(got.mask, want) and got.coverage_fraction == cov


# ------------------------------------------------------------------ masked stretch
@pytest.mark.parametrize("shape,seed", [((64, 64), 1), ((300, 400), 2), ((513, 257), 3)])
def test_masked_stretch_with_mask_bit_exact(pv, orc, shape, seed):
    img = star_field(*shape, 50, seed=seed, sat=True)
    mask, n, cov = orc.generate_star_mask(img, luminance_protect=True)
    mr = pv.StarMaskResult(mask, n, cov)
    got = pv.masked_stretch_with_mask(img, mr)
    want, it, fb, conv = orc.masked_stretch_with_mask(img, mask)
    assert (got.iterations_run, got.final_background, got.converged) == (it, fb, conv)
    assert np.array_equal(got.image, want)
    assert got.iterations_run >= 2 and 0.0 < got.final_background < 1.0


def test_masked_stretch_end_to_end(pv, orc):
    img = star_field(300, 300, 45, seed=9)
    got = pv.masked_stretch(img)
    want, it, fb, conv, n, cov = orc.masked_stretch(img)
    assert (got.iterations_run, got.final_background, got.converged, got.stars_masked,
            got.mask_coverage) == (it, fb, conv, n, cov)
    assert np.array_equal(got.image, want)


def test_masked_stretch_flat_and_empty(pv, orc):
    flat = np.full((20, 30), 5.0, np.float32)  # range 0 -> zeros, medians 0
    mask = np.zeros_like(flat)
    got = pv.masked_stretch_with_mask(flat, pv.StarMaskResult(mask, 0, 0.0))
    want, it, fb, conv = orc.masked_stretch_with_mask(flat, mask)
    assert np.array_equal(got.image, want) and (got.iterations_run, got.final_background) == (it, fb)
    allmasked = np.ones_like(flat)
    img = star_field(20, 30, 3, seed=4)
    got = pv.masked_stretch_with_mask(img, pv.StarMaskResult(allmasked, 0, 1.0))
    want, it, fb, conv = orc.masked_stretch_with_mask(img, allmasked)
    assert np.array_equal(got.image, want) and (got.iterations_run, got.final_background) == (it, fb)


def test_masked_stretch_rgb_shared(pv, orc):
    chans = [star_field(200, 240, 30, seed=21, sky=100.0 + 20 * k) for k in range(3)]
    r, g, b, cov, n = pv.masked_stretch_rgb_shared(*chans)
    lum = orc.synthesize_luminance(*chans, finite_only=True)
    mask, on, ocov = orc.generate_star_mask(lum, luminance_protect=True)
    assert (n, cov) == (on, ocov)
    for res, ch in zip((r, g, b), chans):
        want, it, fb, conv = orc.masked_stretch_with_mask(ch, mask)
        assert np.array_equal(res.image, want) and res.iterations_run == it
