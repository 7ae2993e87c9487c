"""CPU parity of the host-side scalar entry points of libastroburst_cuda.so (no CUDA device needed):
auto-STF parameters, linked-STF stat merge, white-balance reference, histogram downsample, the curve
LUT, the preview size rule and RANSAC -- the library's own host code against the oracle, bit for bit."""
import math

import numpy as np
import pytest


@pytest.fixture(scope="module")
def ab():
    import astroburst_b200

    astroburst_b200.lib()
    return astroburst_b200


def _stats(cls, mn, mx, med, mad, mean=None, n=1000):
    s = cls()
    s.min, s.max, s.median, s.mad = mn, mx, med, mad
    s.sigma = max(mad * 1.4826, 1e-30)
    s.mean = med if mean is None else mean
    s.valid_count = n
    return s


STATS = [(0.0, 65535.0, 1000.0, 8.0), (1.0, 3000.0, 600.0, 180.0), (5.0, 5.0, 5.0, 0.0),
         (0.0, 1.0, 0.999, 0.2), (0.0, 1.0, 1e-9, 0.0), (100.0, 101.0, 100.5, 1e-12)]


def test_auto_stf_linked_stats_wb_reference(ab, orc):
    im = ab.imaging
    for mn, mx, med, mad in STATS:
        g, o = _stats(im.ImageStats, mn, mx, med, mad), _stats(orc.ImageStats, mn, mx, med, mad)
        for tb, k in ((0.25, -2.8), (0.1, -1.25), (0.5, 0.0)):
            a, b = im.auto_stf(g, tb, k), orc.auto_stf(o, tb, k)
            assert (a.shadow, a.midtone, a.highlight) == (b.shadow, b.midtone, b.highlight)
    trip = [STATS[0], STATS[1], STATS[3]]
    gs = [_stats(im.ImageStats, *t, mean=t[2] * 1.01) for t in trip]
    os_ = [_stats(orc.ImageStats, *t, mean=t[2] * 1.01) for t in trip]
    a, b = im.linked_stats(*gs), orc.linked_stats(*os_)
    assert (a.min, a.max, a.median, a.mad, a.sigma, a.mean, a.valid_count) == (
        b.min, b.max, b.median, b.mad, b.sigma, b.mean, b.valid_count)
    assert im.select_wb_reference(*gs) == orc.select_wb_reference(*os_)
    assert im.select_wb_reference(gs[2], gs[2], gs[2]) == orc.select_wb_reference(os_[2], os_[2], os_[2])


def test_downsample_histogram_and_curve_lut(ab, orc):
    im = ab.imaging
    rng = np.random.default_rng(0)
    bins = rng.integers(0, 2 ** 31, 65536, dtype=np.uint32)
    bins[:128] = 2 ** 32 - 1                                   # saturating sums (stats.rs:436-441)
    for target in (512, 100, 65536, 70000, 1):
        assert np.array_equal(im.downsample_histogram(bins, target), orc.downsample_histogram(bins, target))
    for pts in ([], [(0.0, 0.0), (1.0, 1.0)], [(0.0, 0.0), (0.25, 0.15), (0.5, 0.5), (0.75, 0.85), (1.0, 1.0)],
                [(0.3, 0.1), (0.7, 0.9)], [(0.5, 0.5), (0.5, 0.6), (0.2, 0.9)], [(0.1, 0.9), (0.9, 0.1)]):
        assert np.array_equal(im.spline_lut(pts), orc.spline_lut(pts))


def test_preview_size_rule(ab, orc):
    pv = ab.preview
    for rows, cols in ((1, 1), (100, 100), (2048, 2048), (2049, 2048), (13759, 12451), (3, 50000), (50000, 3)):
        for max_dim in (0, 1, 64, 2048):
            assert pv.preview_size(rows, cols, max_dim) == orc.preview_size(rows, cols, max_dim)


def test_ransac_matches_oracle_on_cpu(ab, orc):
    al = ab.alignment
    rng = np.random.default_rng(5)
    ref = rng.uniform(50, 950, (60, 2))
    th = math.radians(2.0)
    A = np.array([[1.01 * math.cos(th), -math.sin(th)], [math.sin(th), 0.99 * math.cos(th)]])
    tgt = ref @ A.T + np.array([20.0, -11.5]) + rng.normal(0, 0.1, ref.shape)
    tgt[:12] = rng.uniform(50, 950, (12, 2))                    # outliers
    m = np.hstack([ref, tgt])
    for method in ("affine", "rigid"):
        for threads in (1, 8, 16):
            a, b = al.ransac_affine(m, method, threads), orc.ransac(m, method, threads)
            assert (a is None) == (b is None)
            if a is not None:
                ta, tb = a.transform, b.transform
                assert (ta.a, ta.b, ta.tx, ta.c, ta.d, ta.ty) == (tb.a, tb.b, tb.tx, tb.c, tb.d, tb.ty)
                assert (a.inliers, a.matched_stars, a.residual_px, a.method) == (
                    b.inliers, b.matched_stars, b.residual_px, b.method)
    assert al.ransac_affine(m[:2], "affine", 8) is None and orc.ransac(m[:2], "affine", 8) is None
