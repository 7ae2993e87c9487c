"""A second, independent restatement of the headline algorithm: sigma_clip_combine and the per-pixel
gather of stack_images (core/stacking/combine.rs:14-92,160-182) written in plain numpy scalar code
straight from the Rust, and a fuzz of the C++ oracle against it.  The oracle is what the CUDA kernels
are compared with, so it gets a cross-check that does not share its code (or its language).

Sums: the reference adds f64 terms in whatever order select_nth_unstable left the samples; the sum
of a few dozen f32 values of comparable magnitude is exact in f64, so any order gives the same
result -- the generators below stay inside that regime (dynamic range far below 2^29), and math.fsum
is used so this file does not depend on an order either."""
import math

import numpy as np

F32 = np.float32


def py_sigma_clip_combine(values, sigma_low, sigma_high, max_iter):
    v = [F32(x) for x in values]
    n = len(v)
    if n == 0:
        return F32(0.0), 0
    if n == 1:
        return v[0], 0
    rejected = 0
    last_center = F32(np.nan)
    sl, sh = F32(sigma_low), F32(sigma_high)
    for it in range(max_iter):
        if len(v) < 2:
            break
        if it == 0:
            s = sorted(v)
            med = s[len(s) // 2]                                     # upper median, not averaged
            devs = sorted(F32(abs(F32(x - med))) for x in v)        # f32 subtraction, f32 abs
            mad = devs[len(devs) // 2]
            center, sigma = med, F32(max(float(mad) * 1.4826, 1e-10))
        else:
            m = len(v)
            mean = math.fsum(float(x) for x in v) / m
            var = math.fsum((float(x) - mean) ** 2 for x in v) / max(m - 1.0, 1.0)
            center, sigma = F32(mean), F32(max(math.sqrt(var), 1e-10))
        last_center = center
        lo, hi = F32(-sl * sigma), F32(sh * sigma)                   # f32 products
        kept = [x for x in v if lo <= F32(x - center) <= hi]         # f32 deviation, inclusive bounds
        removed = len(v) - len(kept)
        rejected += removed
        v = kept
        if removed == 0:
            break
    if not v:
        return (last_center if np.isfinite(last_center) else F32(0.0)), rejected
    return F32(math.fsum(float(x) for x in v) / len(v)), rejected


def test_oracle_sigma_clip_combine_against_python_restatement(orc):
    rng = np.random.default_rng(2024)
    checked = 0
    for trial in range(4000):
        n = int(rng.integers(0, 40))
        base = rng.normal(1000.0, 12.0, n)
        k = int(rng.integers(0, 4))                                  # a few outliers of either sign
        if n and k:
            idx = rng.integers(0, n, k)
            base[idx] += rng.choice([-1.0, 1.0], k) * rng.uniform(50, 40000, k)
        if n > 3 and trial % 7 == 0:
            base[: n // 2] = base[0]                                 # ties: MAD can be 0 -> 1e-10 floor
        vals = base.astype(np.float32)
        sl, sh = [(3.0, 3.0), (2.0, 4.5), (0.5, 0.5), (10.0, 1.0)][trial % 4]
        it = [5, 1, 0, 12][(trial // 4) % 4]
        want, wrej = py_sigma_clip_combine(vals, sl, sh, it)
        got, grej = orc.sigma_clip_combine(vals.copy(), sl, sh, it)
        assert grej == wrej, (trial, vals, sl, sh, it)
        assert F32(got) == want or (np.isnan(got) and np.isnan(want)), (trial, vals, got, want)
        checked += 1
    assert checked == 4000


def test_oracle_stack_images_against_python_restatement(orc):
    rng = np.random.default_rng(7)
    for n_frames, shape in ((1, (3, 4)), (2, (5, 3)), (7, (6, 9)), (32, (4, 8)), (33, (3, 5))):
        frames = rng.normal(500.0, 6.0, (n_frames,) + shape).astype(np.float32)
        hit = rng.random(frames.shape) < 0.03
        frames[hit] += rng.uniform(200, 9000, int(hit.sum())).astype(np.float32)
        bad = rng.random(frames.shape) < 0.04                        # combine.rs:170-175 drops non-finite
        frames[bad] = rng.choice(np.array([np.nan, np.inf, -np.inf], np.float32), int(bad.sum()))
        if n_frames > 2:
            frames[:, 0, 0] = np.nan                                 # a pixel with no finite sample -> 0
        want = np.zeros(shape, np.float32)
        wrej = 0
        for r in range(shape[0]):
            for c in range(shape[1]):
                col = [x for x in frames[:, r, c] if np.isfinite(x)]
                want[r, c], k = py_sigma_clip_combine(col, 3.0, 3.0, 5)
                wrej += k
        got, grej = orc.stack_images(list(frames), 3.0, 3.0, 5)
        assert grej == wrej and np.array_equal(got, want), (n_frames, shape)


# ---- image statistics, histogram path (core/imaging/stats.rs:75-210, helpers :212-353), restated with
# numpy array operations (bincount / cumsum / searchsorted) instead of the oracle's scalar loops
BINS = 65536


def _as_usize(x):
    """Rust `f64 as usize` for arrays: truncation toward zero, negatives and NaN -> 0."""
    x = np.nan_to_num(x, nan=0.0)
    return np.clip(x, 0.0, 1.8e19).astype(np.uint64)


def _pct_bin(hist, total, pct):
    target = np.uint64(math.ceil(total * pct))
    cum = np.cumsum(hist, dtype=np.uint64)
    i = int(np.searchsorted(cum, target, side="left"))               # first cum >= target
    return (i if i < len(hist) else len(hist) - 1), cum, target


def _interp(hist, total, pct, lo, width):
    target = np.uint64(math.ceil(total * pct))
    cum = np.cumsum(hist, dtype=np.uint64)
    i = int(np.searchsorted(cum, target, side="left"))
    if i >= len(hist):
        return lo + len(hist) * width
    over = int(cum[i] - target)
    frac = 1.0 - over / float(hist[i]) if hist[i] > 0 else 0.5
    return lo + (i + frac) * width


def _rank(hist, rank, lo, width):
    if rank == 0:
        return lo
    cum = np.cumsum(hist, dtype=np.uint64)
    i = int(np.searchsorted(cum, np.uint64(rank), side="left"))
    if i >= len(hist):
        return lo + len(hist) * width
    over = int(cum[i]) - int(rank)
    frac = 1.0 - over / float(hist[i]) if hist[i] > 0 else 0.5
    return lo + (i + frac) * width


def np_stats_hist(data, known=None):
    v = np.asarray(data, np.float32).reshape(-1)
    v = v[np.isfinite(v) & (v > np.float32(1e-7))]
    if v.size == 0:
        return None
    vf = v.astype(np.float64)
    gmin, gmax = (float(vf.min()), float(vf.max())) if known is None else known
    rng_ = max(gmax - gmin, 1e-30)
    bw, inv = rng_ / BINS, BINS / rng_
    hist = np.bincount(np.minimum(_as_usize((vf - gmin) * inv), BINS - 1).astype(np.int64), minlength=BINS)
    total = int(v.size)
    half = int(math.ceil(total * 0.5))
    mbin, cum, _ = _pct_bin(hist, total, 0.5)
    before = int(cum[mbin - 1]) if mbin > 0 else 0
    mlo = gmin + mbin * bw
    mhi = mlo + bw
    coarse = _interp(hist, total, 0.5, gmin, bw)
    rrange = max(mhi - mlo, 1e-30)
    rinv = BINS / rrange
    inbin = (vf >= mlo) & (vf < mhi)
    refine = np.bincount(np.minimum(_as_usize((vf[inbin] - mlo) * rinv), BINS - 1).astype(np.int64), minlength=BINS)
    d = np.abs(v - np.float32(coarse))                               # f32 subtraction
    dev = np.bincount(np.minimum(_as_usize(d.astype(np.float64) * inv), BINS - 1).astype(np.int64), minlength=BINS)
    median = _rank(refine, max(half - before, 0), mlo, rrange / BINS)
    madbin, _, _ = _pct_bin(dev, total, 0.5)
    elo, ehi = max(madbin - 1, 0), min(madbin + 2, BINS)
    rlo, rhi = elo * bw, ehi * bw
    mrange = max(rhi - rlo, 1e-30)
    minv = BINS / mrange
    dd = np.abs(v - np.float32(median))
    below = int((dd < np.float32(rlo)).sum())
    sel = (dd >= np.float32(rlo)) & (dd < np.float32(rhi))
    mref = np.bincount(np.minimum(_as_usize((dd[sel].astype(np.float64) - rlo) * minv), BINS - 1).astype(np.int64),
                       minlength=BINS)
    mad = _rank(mref, max(half - below, 0), rlo, mrange / BINS)
    return dict(min=gmin, max=gmax, median=median, mad=mad, sigma=max(mad * 1.4826, 1e-30), n=total,
                mean=math.fsum(vf.tolist()) / total)


def test_oracle_hist_stats_against_numpy_restatement(orc):
    rng = np.random.default_rng(11)
    cases = []
    sky = rng.normal(1000.0, 12.0, (300, 400))
    sky.reshape(-1)[rng.integers(0, sky.size, 300)] += rng.uniform(100, 40000, 300)
    sky[:8] = 0.0
    sky[20, 20] = np.nan
    cases.append(sky.astype(np.float32))
    cases.append(rng.uniform(1.0, 3000.0, (257, 129)).astype(np.float32))
    cases.append(np.full((64, 64), 5.0, np.float32))                 # zero range
    two = np.where(rng.random((100, 100)) < 0.5, 10.0, 20.0).astype(np.float32)
    cases.append(two)                                                # two-valued: empty median bins
    cases.append((rng.exponential(50.0, (200, 200)) + 1e-3).astype(np.float32))
    for img in cases:
        want = np_stats_hist(img)
        got = orc.compute_image_stats_hist(img)
        assert (got.min, got.max, got.valid_count) == (want["min"], want["max"], want["n"])
        assert got.median == want["median"] and got.mad == want["mad"] and got.sigma == want["sigma"]
        assert abs(got.mean - want["mean"]) <= 1e-12 * abs(want["mean"])
        kr = (float(want["min"]) - 3.0, float(want["max"]) * 1.5)
        w2 = np_stats_hist(img, known=kr)
        g2 = orc.compute_image_stats_with_known_range(np.tile(img, (8, 8)), *kr) if img.size * 64 > 4000000 else None
        if g2 is not None:                                           # the known-range form switches at 4 M px
            t2 = np_stats_hist(np.tile(img, (8, 8)), known=kr)
            assert (g2.median, g2.mad, g2.valid_count) == (t2["median"], t2["mad"], t2["n"])
        assert w2 is not None
    assert orc.compute_image_stats_hist(np.zeros((10, 10), np.float32)).valid_count == 0


# ---- tile background and star detection (core/analysis/star_detection.rs:32-258, math/sigma_clip.rs,
# math/median.rs) restated with numpy medians and scipy connected components.  Moment sums run in
# raster order here and in BFS order in the reference/oracle, so star parameters are compared to
# 1e-12 relative; counts, npix, ordering and the background are exact.
def np_sigma_clipped_stats(vals, kappa=3.0, iterations=2):
    v = np.asarray(vals, np.float32)

    def med64(a):                                                    # exact_median_mut: even n -> f64 mean of the two middle f32
        s = np.sort(a)
        n = s.size
        return float(s[n // 2]) if n % 2 else (float(s[n // 2 - 1]) + float(s[n // 2])) / 2.0

    def mad32(a, median):                                            # devs cast to f32, median_f32_mut averages in f32
        d = np.sort(np.abs(a.astype(np.float64) - median).astype(np.float32))
        n = d.size
        return float(d[n // 2]) if n % 2 else float(np.float32((d[n // 2 - 1] + d[n // 2]) / np.float32(2.0)))

    for _ in range(iterations):
        if v.size < 3:
            break
        median = med64(v)
        sig = max(mad32(v, median) * 1.4826, 1e-30)
        lo, hi = np.float32(median - kappa * sig), np.float32(median + kappa * sig)
        v = v[(v >= lo) & (v <= hi)]
    if v.size == 0:
        return 0.0, 1.0
    median = med64(v)
    return median, max(mad32(v, median) * 1.4826, 1e-30)


def np_estimate_background(img, tile):
    rows, cols = img.shape
    step = max(tile, 16)
    meds, sigs = [], []
    for y in range(0, rows, step):
        for x in range(0, cols, step):
            t = img[y:y + step, x:x + step].reshape(-1)
            t = t[np.isfinite(t) & (t > np.float32(1e-7))]
            if t.size >= 8:
                m, s = np_sigma_clipped_stats(t)
                meds.append(m)
                sigs.append(s)
    if not meds:
        return 0.0, 1.0
    meds.sort()
    sigs.sort()
    return meds[len(meds) // 2], max(sigs[len(sigs) // 2], 1e-10)


def np_detect_stars(img, sigma_threshold):
    from scipy import ndimage

    rows, cols = img.shape
    tile = min(max(min(rows, cols) // 8, 32), 256)
    bg, sg = np_estimate_background(img, tile)
    thr = bg + sigma_threshold * sg
    fg = np.isfinite(img) & (img.astype(np.float64) > thr)
    lab, n = ndimage.label(fg, structure=np.ones((3, 3), int))
    stars = []
    for k in range(1, n + 1):
        rr, cc = np.nonzero(lab == k)                                # raster order
        interior = (rr >= 1) & (rr < rows - 1) & (cc >= 1) & (cc < cols - 1)
        if not interior.any():                                       # never seeded (seeds are interior pixels)
            continue
        seed = int(np.flatnonzero(interior)[0])
        npix = rr.size
        if npix < 3 or npix > 5000:
            continue
        w = np.maximum(img[rr, cc].astype(np.float64) - bg, 0.0)
        flux = float(w.sum())
        if flux <= 0.0:
            continue
        cx, cy = float((cc * w).sum() / flux), float((rr * w).sum() / flux)
        dx, dy = cc - cx, rr - cy
        fwhm = math.sqrt(float(((dx * dx + dy * dy) * w).sum()) / (2.0 * flux)) * 2.3548200450309493
        if fwhm < 0.5 or fwhm > 30.0:
            continue
        ixx, iyy, ixy = float((dx * dx * w).sum() / flux), float((dy * dy * w).sum() / flux), float((dx * dy * w).sum() / flux)
        tr = ixx + iyy
        det = max(ixx * iyy - ixy * ixy, 0.0)
        disc = math.sqrt(max(tr * tr / 4.0 - det, 0.0))
        l1, l2 = tr / 2.0 + disc, max(tr / 2.0 - disc, 0.0)
        ecc = min(max(math.sqrt(1.0 - l2 / l1), 0.0), 1.0) if l1 > 1e-15 else 0.0
        peak = float(w.max())
        order = (int(rr[seed]), int(cc[seed]))                       # discovery order = raster order of seeds
        stars.append((order, dict(x=cx, y=cy, flux=flux, fwhm=fwhm, eccentricity=ecc, peak=peak, npix=int(npix),
                                  snr=peak / sg if sg > 2.220446049250313e-16 else 0.0)))
    stars.sort(key=lambda t: t[0])
    stars = [s for _, s in stars]
    stars.sort(key=lambda s: -s["flux"])                             # stable, like sort_by
    kept = []
    for s in stars:                                                  # greedy 3-px dedup (any kept star within 3 px)
        if all((s["x"] - t["x"]) ** 2 + (s["y"] - t["y"]) ** 2 >= 9.0 for t in kept):
            kept.append(s)
    return kept, bg, sg


def test_oracle_background_and_detection_against_numpy_scipy_restatement(orc):
    rng = np.random.default_rng(3)
    for shape, n_stars in (((200, 260), 40), ((333, 257), 90)):
        img = rng.normal(100.0, 3.0, shape)
        for _ in range(n_stars):
            x, y = rng.uniform(-3, shape[1] + 3), rng.uniform(-3, shape[0] + 3)
            f = 300.0 * (100.0 ** rng.uniform(0, 1))
            yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
            img += f / (2 * np.pi * 1.6 ** 2) * np.exp(-((xx - x) ** 2 + (yy - y) ** 2) / (2 * 1.6 ** 2))
        img = img.astype(np.float32)
        img[:6, :] = 0.0                                             # padding, excluded from the background
        img[50, 60] = np.nan
        img[70:74, 80:86] += 4000.0                                  # a blob touching nothing
        img[120:123, 0:2] += 4000.0                                  # a component on the image edge
        for tile in (32, 64):
            assert orc.estimate_background(img, tile) == np_estimate_background(img, tile)
        want, bg, sg = np_detect_stars(img, 4.0)
        got = orc.detect_stars(img, 4.0)
        assert (got.background_median, got.background_sigma) == (bg, sg)
        assert len(got.stars) == len(want) and len(want) > 10
        for g, w in zip(got.stars, want):
            assert g["npix"] == w["npix"]
            for key in ("x", "y", "flux", "fwhm", "peak", "snr"):
                assert abs(g[key] - w[key]) <= 1e-12 * max(abs(w[key]), 1.0), key
            assert abs(g["eccentricity"] - w["eccentricity"]) <= 1e-9


# ---- Catmull-Rom bicubic gathers (imaging/sampling.rs:4-13,51-80, boundary.rs:8-20) and their three
# callers (alignment/affine.rs:663-690 warp_image, stacking/align.rs:36-57 shift_image_subpixel,
# imaging/resample.rs:25-61 resample_image), restated as whole-array numpy f64 arithmetic in the
# reference's evaluation order: bit-identical pixels expected.
def _cr(t):
    a = np.abs(t)
    return np.where(a <= 1.0, a * a * (1.5 * a - 2.5) + 1.0,
                    np.where(a <= 2.0, a * (a * (2.5 - 0.5 * a) - 4.0) + 2.0, 0.0))


def np_bicubic(img, sy, sx):
    rows, cols = img.shape
    ix, iy = np.floor(sx), np.floor(sy)
    fx, fy = sx - ix, sy - iy
    ix, iy = ix.astype(np.int64), iy.astype(np.int64)
    wx = [_cr(fx + 1.0), _cr(fx), _cr(fx - 1.0), _cr(fx - 2.0)]
    val = np.zeros(sx.shape, np.float64)
    for j in range(4):
        r = np.clip(iy + j - 1, 0, rows - 1)
        row_val = np.zeros(sx.shape, np.float64)
        for i in range(4):
            c = np.clip(ix + i - 1, 0, cols - 1)
            row_val = row_val + img[r, c].astype(np.float64) * wx[i]
        val = val + row_val * _cr(fy - float(j - 1))
    return val.astype(np.float32)


def np_warp(img, t, out_rows, out_cols):
    a, b, tx, c, d, ty = t
    y, x = np.mgrid[0:out_rows, 0:out_cols].astype(np.float64)
    sx, sy = a * x + b * y + tx, c * x + d * y + ty
    ok = (sx >= 0.0) & (sy >= 0.0) & (sx < img.shape[1] - 1) & (sy < img.shape[0] - 1)
    out = np.zeros((out_rows, out_cols), np.float32)
    out[ok] = np_bicubic(img, sy[ok], sx[ok])
    return out


def np_shift(img, dy, dx):
    if abs(dy) < 1e-12 and abs(dx) < 1e-12:
        return img.copy()
    rows, cols = img.shape
    y, x = np.mgrid[0:rows, 0:cols].astype(np.float64)
    sy, sx = y + dy, x + dx
    skip = (sy < -0.5) | (sy > rows - 0.5) | (sx < -0.5) | (sx > cols - 0.5)
    out = np.zeros(img.shape, np.float32)
    out[~skip] = np_bicubic(img, sy[~skip], sx[~skip])
    return out


def np_resample(img, tr, tc):
    if (tr, tc) == img.shape:
        return img.copy()
    sy_, sx_ = img.shape[0] / tr, img.shape[1] / tc
    ty, tx = np.mgrid[0:tr, 0:tc].astype(np.float64)
    return np_bicubic(img, ty * sy_ + (sy_ - 1.0) * 0.5, tx * sx_ + (sx_ - 1.0) * 0.5)


def test_oracle_bicubic_gathers_against_numpy_restatement(orc):
    rng = np.random.default_rng(9)
    img = rng.normal(500.0, 40.0, (97, 131)).astype(np.float32)
    th = math.radians(2.3)
    for t in ((1, 0, 0, 0, 1, 0), (1, 0, 4.5, 0, 1, -2.25),
              (1.01 * math.cos(th), -math.sin(th), 7.3, math.sin(th), 0.99 * math.cos(th), -3.9),
              (0.5, 0, 0.25, 0, 0.5, 0.75), (1, 0, 500.0, 0, 1, 500.0)):
        for shape in ((97, 131), (60, 200)):
            assert np.array_equal(orc.warp_image(img, t, *shape), np_warp(img, t, *shape)), (t, shape)
    for dy, dx in ((0.0, 0.0), (0.5, -0.5), (3.25, 7.75), (-120.0, 2.0), (0.3, 130.6), (1e-13, 0.0)):
        assert np.array_equal(orc.shift_image_subpixel(img, dy, dx), np_shift(img, dy, dx)), (dy, dx)
    for tr, tc in ((97, 131), (48, 65), (194, 262), (33, 400), (1, 1)):
        assert np.array_equal(orc.resample_image(img, tr, tc), np_resample(img, tr, tc)), (tr, tc)


# ---- triangle matching (core/alignment/affine.rs:279-398) restated with numpy broadcasting.  The
# reference sorts the vote pairs by votes only (HashMap order breaks ties); the oracle and the CUDA
# path break ties by (ref, tgt) index -- the restatement does the same, so lists compare exactly.
def np_match_stars(ref, tgt, tol=0.02, min_side=15.0, min_votes=1):
    def tris(p):
        p = np.asarray(p, np.float64)[:60]
        n = len(p)
        out = []
        for i in range(n):
            for j in range(i + 1, n):
                for k in range(j + 1, n):
                    d = lambda a, b: math.sqrt((a[0] - b[0]) ** 2 + (a[1] - b[1]) ** 2)  # affine.rs:659-661
                    dij, djk, dik = d(p[i], p[j]), d(p[j], p[k]), d(p[i], p[k])
                    s = sorted((dij, djk, dik))
                    if s[0] < min_side:
                        continue
                    # vertices ordered by the length of the OPPOSITE side (stable)
                    verts = sorted(((i, djk), (j, dik), (k, dij)), key=lambda t: t[1])
                    out.append((s[1] / s[0], s[2] / s[0], [v[0] for v in verts]))
        return out

    rt, tt = tris(ref), tris(tgt)
    if not rt or not tt:
        return np.zeros((0, 4))
    rm, rl = np.array([t[0] for t in rt]), np.array([t[1] for t in rt])
    tm, tl = np.array([t[0] for t in tt]), np.array([t[1] for t in tt])
    rv, tv = np.array([t[2] for t in rt]), np.array([t[2] for t in tt])
    votes = np.zeros((len(ref), len(tgt)), np.int64)
    for a in range(len(rt)):
        hit = np.flatnonzero(~((np.abs(rm[a] - tm) > tol) | (np.abs(rl[a] - tl) > tol)))
        for p in range(3):
            np.add.at(votes, (rv[a, p], tv[hit, p]), 1)
    pairs = [(int(votes[r, t]), r, t) for r in range(votes.shape[0]) for t in range(votes.shape[1]) if votes[r, t]]
    pairs.sort(key=lambda x: (-x[0], x[1], x[2]))
    used_r, used_t, out = set(), set(), []
    for v, r, t in pairs:
        if v < min_votes:
            break
        if r in used_r or t in used_t:
            continue
        used_r.add(r)
        used_t.add(t)
        out.append((ref[r][0], ref[r][1], tgt[t][0], tgt[t][1]))
    return np.array(out, np.float64).reshape(-1, 4)


def test_oracle_triangle_matching_against_numpy_restatement(orc):
    rng = np.random.default_rng(17)
    ref = rng.uniform(40, 960, (18, 2))
    th = math.radians(3.0)
    A = np.array([[1.02 * math.cos(th), -math.sin(th)], [math.sin(th), 0.98 * math.cos(th)]])
    tgt = (ref @ A.T + np.array([15.0, -9.0]) + rng.normal(0, 0.05, ref.shape))[rng.permutation(18)][:15]
    tgt = np.vstack([tgt, rng.uniform(40, 960, (3, 2))])             # strays
    want = np_match_stars(ref, tgt)
    got = orc.match_stars(ref, tgt, max_matches=256)
    assert len(want) >= 10 and got.shape == want.shape and np.array_equal(got, want)
    assert orc.match_stars(ref[:2], tgt).shape[0] == 0              # fewer than 3 stars: no triangles


# ---- the colour chain's pointwise operators as whole-array numpy arithmetic in the reference's
# evaluation order and precision (f64 for STF, f32 for blend / SCNR): bit-identical results expected.
def np_auto_stf(st, target_bg=0.25, shadow_k=-2.8):                  # stf.rs:13-47
    if st.valid_count == 0:
        return 0.0, 0.5, 1.0
    rng_ = max(st.max - st.min, 1e-30)
    med_n, sig_n = (st.median - st.min) / rng_, st.sigma / rng_
    shadow = min(max(med_n + shadow_k * sig_n, 0.0), 0.98)
    clip = max(1.0 - shadow, 1e-15)
    m = min(max((med_n - shadow) / clip, 0.0), 1.0)
    if m <= 0.0 or m >= 1.0:
        mid = 0.5
    else:
        den = 2.0 * target_bg * m - target_bg - m
        mid = 0.5 if abs(den) < 1e-15 else min(max(m * (target_bg - 1.0) / den, 0.0001), 0.9999)
    return shadow, mid, 1.0


def np_apply_stf(v, stf, st, as_u8):                                 # stf.rs:50-120
    shadow, mid, hi = stf
    valid = np.isfinite(v) & (v > np.float32(1e-7))
    x = v.astype(np.float64)
    norm = (x - st.min) * (1.0 / max(st.max - st.min, 1e-30))
    c = np.clip((norm - shadow) * (1.0 / max(hi - shadow, 1e-15)), 0.0, 1.0)
    with np.errstate(all="ignore"):
        y = np.where(c <= 0.0, 0.0, np.where(c >= 1.0, 1.0, (mid - 1.0) * c / ((2.0 * mid - 1.0) * c - mid)))
    if as_u8:
        q = y * 255.0
        q = np.clip(np.sign(q) * np.floor(np.abs(q) + 0.5), 0.0, 255.0)   # f64::round: half away from zero
        return np.where(valid, np.nan_to_num(q), 0.0).astype(np.uint8)
    return np.where(valid, y, 0.0).astype(np.float32)


def np_scnr(r, g, b, method, amount, preserve):                      # scnr.rs:5-53
    amount = np.float32(min(max(amount, 0.0), 1.0))
    if amount < np.float32(1e-7):
        return r, g, b
    limit = (r + b) * np.float32(0.5) if method == "average" else np.maximum(r, b)
    g_new = g + amount * (np.minimum(g, limit) - g)
    delta = g - g_new
    inv_rb = np.float32(1.0) / (np.float32(0.2126) + np.float32(0.0722))
    boost = (np.float32(0.7152) * delta) * inv_rb
    do = preserve & (delta > np.float32(1e-10)) & (r <= 1.0) & (b <= 1.0)
    r2 = np.where(do, np.minimum(r + boost, np.float32(1.0)), r)
    b2 = np.where(do, np.minimum(b + boost, np.float32(1.0)), b)
    return r2.astype(np.float32), g_new.astype(np.float32), b2.astype(np.float32)


def test_oracle_colour_chain_against_numpy_restatement(orc):
    rng = np.random.default_rng(23)
    chans = []
    for k in range(3):
        c = rng.normal(900.0 + 120 * k, 25.0, (120, 150))
        c.reshape(-1)[rng.integers(0, c.size, 60)] += rng.uniform(200, 30000, 60)
        c = c.astype(np.float32)
        c[:4] = 0.0
        c[10, 10], c[11, 11], c[12, 12] = np.nan, np.inf, -7.0
        chans.append(c)
    # blend: rv += v * w in weight order, f32, unfused (channel_blend.rs:45-63)
    w = [(0, 0.8, 0.1, 0.0), (1, 0.2, 0.7, 0.15), (2, 0.0, 0.2, 0.85), (5, 1.0, 1.0, 1.0)]
    want = [np.zeros((120, 150), np.float32) for _ in range(3)]
    for ci, rw, gw, bw in w:
        if ci >= len(chans):
            continue
        with np.errstate(invalid="ignore"):                      # inf * 0 weight -> NaN, as in the reference
            for o, wt in zip(want, (rw, gw, bw)):
                o += chans[ci] * np.float32(wt)
    got = orc.blend_channels(chans, w, 120, 150)
    for x, y in zip(got, want):
        assert np.array_equal(x, y, equal_nan=True)
    # stats -> auto-STF -> u8 and f32 stretch
    for c in chans:
        st = orc.compute_image_stats(c)
        p = orc.auto_stf(st)
        assert (p.shadow, p.midtone, p.highlight) == np_auto_stf(st)
        stf = (p.shadow, p.midtone, p.highlight)
        assert np.array_equal(orc.apply_stf(c, p, st), np_apply_stf(c, stf, st, True))
        assert np.array_equal(orc.apply_stf_f32(c, p, st), np_apply_stf(c, stf, st, False))
    # SCNR on stretched planes
    st = [orc.compute_image_stats(c) for c in chans]
    planes = [orc.apply_stf_f32(c, orc.auto_stf(s), s) for c, s in zip(chans, st)]
    planes[0][5, 5] = 1.5                                             # r > 1 blocks the luminance boost
    for method in ("average", "maximum"):
        for amount, preserve in ((1.0, False), (0.5, True), (1e-9, True), (3.0, True)):
            a = [p.copy() for p in planes]
            orc.apply_scnr_inplace(*a, method, amount, preserve)
            for x, y in zip(a, np_scnr(*planes, method, amount, preserve)):
                assert np.array_equal(x, y), (method, amount, preserve)


# ---- exact statistics path (stats.rs:43-73, median.rs:27-73) and build/downsample_histogram
# (stats.rs:378-444) with numpy sort / bincount / reduceat.
def test_oracle_exact_stats_and_histogram_against_numpy_restatement(orc):
    rng = np.random.default_rng(31)
    for n in (1, 2, 7, 1000, 1001):
        v = rng.normal(800.0, 30.0, n).astype(np.float32)
        if n > 5:
            v[:2] = [0.0, np.nan]
        ok = v[np.isfinite(v) & (v > np.float32(1e-7))]
        got = orc.compute_image_stats_exact(v)
        if ok.size == 0:
            assert got.valid_count == 0
            continue
        s = np.sort(ok)
        m = s.size // 2
        med = float(s[m]) if s.size % 2 else (float(s[m - 1]) + float(s[m])) / 2.0
        d = np.sort(np.abs(ok - np.float32(med)))                    # f32 deviations about the f32 median
        mad = float(d[m]) if d.size % 2 else float(np.float32((d[m - 1] + d[m]) / np.float32(2.0)))
        assert (got.min, got.max, got.median, got.mad, got.valid_count) == (float(s[0]), float(s[-1]), med, mad, s.size)
        assert got.sigma == max(mad * 1.4826, 1e-30)
        assert abs(got.mean - math.fsum(ok.astype(np.float64).tolist()) / s.size) <= 1e-12 * abs(got.mean)
    img = rng.normal(1000.0, 15.0, (200, 300)).astype(np.float32)
    img[:3] = 0.0
    img[9, 9] = np.inf
    ok = img[np.isfinite(img) & (img > np.float32(1e-7))].astype(np.float64)
    lo, hi = float(ok.min()), float(ok.max())
    bins, edges = orc.build_histogram(img, 65536, lo, hi)
    idx = np.minimum(((ok - lo) * (65536 / (hi - lo))).astype(np.int64), 65535)
    assert np.array_equal(bins, np.bincount(idx, minlength=65536).astype(np.uint32))
    assert np.array_equal(edges, lo + np.arange(65537, dtype=np.float64) * ((hi - lo) / 65536))
    z, ez = orc.build_histogram(img, 65536, 5.0, 5.0)                # range < 1e-10: zeros, flat edges
    assert not z.any() and np.all(ez == 5.0)
    for target in (512, 100, 3):
        ratio = 65536 / target
        starts = (np.arange(target) * ratio).astype(np.int64)
        ends = np.minimum(((np.arange(target) + 1) * ratio).astype(np.int64), 65536)
        want = np.array([min(int(bins[a:b].astype(np.uint64).sum()), 2 ** 32 - 1) for a, b in zip(starts, ends)], np.uint32)
        assert np.array_equal(orc.downsample_histogram(bins, target), want)


# ---- levels and the monotone-spline tone curve (core/imaging/curves.rs:17-208) in scalar Python
def py_spline_lut(points):
    pts = sorted(((float(x), float(y)) for x, y in points), key=lambda p: p[0])      # stable, like sort_by
    ded = []
    for p in pts:                                                    # dedup_by keeps the first of a run
        if ded and abs(p[0] - ded[-1][0]) < 1e-9:
            continue
        ded.append(p)
    pts = ded
    if not pts or pts[0][0] > 1e-6:
        pts.insert(0, (0.0, 0.0))
    if pts[-1][0] < 1.0 - 1e-6:
        pts.append((1.0, 1.0))
    n = len(pts)
    sgn = lambda v: math.copysign(1.0, v)                            # f64::signum: +0 -> 1, -0 -> -1
    if n == 2:
        m = [(pts[1][1] - pts[0][1]) / max(pts[1][0] - pts[0][0], 1e-15)] * 2
    else:
        dx = [max(pts[i + 1][0] - pts[i][0], 1e-15) for i in range(n - 1)]
        sl = [(pts[i + 1][1] - pts[i][1]) / dx[i] for i in range(n - 1)]
        m = [0.0] * n
        m[0], m[-1] = sl[0], sl[-1]
        for i in range(1, n - 1):
            m[i] = 0.0 if sgn(sl[i - 1]) != sgn(sl[i]) else (sl[i - 1] + sl[i]) * 0.5
        for i in range(n - 1):
            if abs(sl[i]) < 1e-15:
                m[i] = m[i + 1] = 0.0
                continue
            a, b = m[i] / sl[i], m[i + 1] / sl[i]
            tau = a * a + b * b
            if tau > 9.0:
                s = 3.0 / math.sqrt(tau)
                m[i], m[i + 1] = s * a * sl[i], s * b * sl[i]

    def ev(x):
        if x <= pts[0][0]:
            return pts[0][1]
        if x >= pts[-1][0]:
            return pts[-1][1]
        seg = 0
        for i in range(1, n):
            if x < pts[i][0]:
                seg = i - 1
                break
        d = max(pts[seg + 1][0] - pts[seg][0], 1e-15)
        t = (x - pts[seg][0]) / d
        t2 = t * t
        t3 = t2 * t
        return ((2.0 * t3 - 3.0 * t2 + 1.0) * pts[seg][1] + (t3 - 2.0 * t2 + t) * d * m[seg]
                + (-2.0 * t3 + 3.0 * t2) * pts[seg + 1][1] + (t3 - t2) * d * m[seg + 1])

    return np.array([np.float32(min(max(ev(i / 4095.0), 0.0), 1.0)) for i in range(4096)], np.float32)


def test_oracle_levels_and_curves_against_python_restatement(orc):
    for pts in ([], [(0.0, 0.0), (1.0, 1.0)], [(0.0, 0.0), (0.25, 0.15), (0.5, 0.5), (0.75, 0.85), (1.0, 1.0)],
                [(0.3, 0.1), (0.7, 0.9)], [(0.5, 0.5), (0.5, 0.6), (0.2, 0.9)], [(0.1, 0.9), (0.9, 0.1)],
                [(0.0, 0.2), (0.4, 0.2), (0.6, 0.8), (1.0, 0.8)]):
        assert np.array_equal(orc.spline_lut(pts), py_spline_lut(pts)), pts
    rng = np.random.default_rng(41)
    u = rng.uniform(-0.2, 1.4, (90, 110)).astype(np.float32)
    u[0, 0], u[0, 1] = np.nan, np.inf
    lut = orc.spline_lut([(0.0, 0.0), (0.25, 0.15), (0.5, 0.5), (0.75, 0.85), (1.0, 1.0)])
    bad = ~np.isfinite(u) | (u < 0)
    idx = np.minimum((np.clip(np.nan_to_num(u), 0.0, 1.0).astype(np.float32) * np.float32(4095.0)).astype(np.int64), 4095)
    assert np.array_equal(orc.apply_curve(u, lut), np.where(bad, np.float32(0.0), lut[idx]))
    for black, gamma, white in ((0.0, 1.0, 1.0), (0.2, 1.0, 1.0), (0.05, 2.0, 0.9), (0.1, 100.0, 0.1), (0.0, 0.001, 1.0)):
        got = orc.apply_levels(u, black, gamma, white)
        if abs(black) < 1e-7 and abs(gamma - 1.0) < 1e-7 and abs(white - 1.0) < 1e-7:
            assert np.array_equal(got, u, equal_nan=True)            # identity short-circuit returns a clone
            continue
        inv_r, inv_g = 1.0 / max(white - black, 1e-15), 1.0 / min(max(gamma, 0.01), 10.0)
        norm = np.clip((np.nan_to_num(u).astype(np.float64) - black) * inv_r, 0.0, 1.0)
        want = np.where(bad, np.float32(0.0), np.array([math.pow(x, inv_g) for x in norm.reshape(-1)]).reshape(u.shape).astype(np.float32))
        assert np.array_equal(got, want)


# ---- RANSAC and the two fits (core/alignment/affine.rs:400-657) in scalar Python: xorshift64 streams
# seeded per emulated rayon thread, least-squares affine via Cramer's rule, closed-form rigid fit.
M64 = (1 << 64) - 1


def py_fit_affine(ms):
    if len(ms) < 3:
        return None

    def solve(which):
        ata = [[0.0] * 3 for _ in range(3)]
        atb = [0.0] * 3
        for rx, ry, tx, ty in ms:
            row, target = (rx, ry, 1.0), (tx if which == 0 else ty)
            for i in range(3):
                for j in range(3):
                    ata[i][j] += row[i] * row[j]
                atb[i] += row[i] * target
        a = ata
        det = (a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0])
               + a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]))
        if abs(det) < 1e-12:
            return None
        d = 1.0 / det
        inv = [[(a[1][1] * a[2][2] - a[1][2] * a[2][1]) * d, (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * d,
                (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * d],
               [(a[1][2] * a[2][0] - a[1][0] * a[2][2]) * d, (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * d,
                (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * d],
               [(a[1][0] * a[2][1] - a[1][1] * a[2][0]) * d, (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * d,
                (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * d]]
        return [inv[i][0] * atb[0] + inv[i][1] * atb[1] + inv[i][2] * atb[2] for i in range(3)]

    x, y = solve(0), solve(1)
    if x is None or y is None:
        return None
    return (x[0], x[1], x[2], y[0], y[1], y[2])


def py_fit_rigid(ms):
    n = len(ms)
    if n < 2:
        return None
    rcx = rcy = tcx = tcy = 0.0
    for rx, ry, tx, ty in ms:
        rcx += rx
        rcy += ry
        tcx += tx
        tcy += ty
    rcx, rcy, tcx, tcy = rcx / n, rcy / n, tcx / n, tcy / n
    num = den = 0.0
    for rx, ry, tx, ty in ms:
        drx, dry, dtx, dty = rx - rcx, ry - rcy, tx - tcx, ty - tcy
        num += drx * dty - dry * dtx
        den += drx * dtx + dry * dty
    th = math.atan2(num, den)
    c, s = math.cos(th), math.sin(th)
    return (c, -s, tcx - c * rcx + s * rcy, s, c, tcy - s * rcx - c * rcy)


def py_ransac(ms, method, n_threads):
    n = len(ms)
    k = 3 if method == "affine" else 2
    if n < k:
        return None
    fit = py_fit_affine if method == "affine" else py_fit_rigid
    chunk = (2000 + n_threads - 1) // n_threads
    best = (0, (1.0, 0.0, 0.0, 0.0, 1.0, 0.0), [False] * n)
    for tid in range(n_threads):
        state = (0xDEADBEEFCAFEBABE + tid * 0x9E3779B97F4A7C15) & M64
        lb, lt, lm = 0, best[1], [False] * n
        for _ in range(chunk):
            sample, attempts = [], 0
            while len(sample) < k and attempts < 20:
                state ^= (state << 13) & M64
                state ^= state >> 7
                state ^= (state << 17) & M64
                idx = state % n
                if idx not in sample:
                    sample.append(idx)
                attempts += 1
            if len(sample) < k:
                continue
            t = fit([ms[i] for i in sample])
            if t is None:
                continue
            mask = [math.sqrt((t[0] * rx + t[1] * ry + t[2] - tx) ** 2 + (t[3] * rx + t[4] * ry + t[5] - ty) ** 2) < 3.0
                    for rx, ry, tx, ty in ms]
            inl = sum(mask)
            if inl > lb:
                lb, lt, lm = inl, t, mask
        if tid == 0 or lb > best[0]:                                 # reduce_with keeps the first maximum
            best = (lb, lt, lm)
    bi, bt, bm = best
    if bi < 4 or bi / n < 0.20:
        return None
    inl = [m for m, keep in zip(ms, bm) if keep]
    ref = fit(inl) or bt
    res = 0.0                                                        # plain sequential f64 adds (Python >= 3.12's
    for rx, ry, tx, ty in inl:                                       # sum() is compensated, the reference's is not)
        res += math.sqrt((ref[0] * rx + ref[1] * ry + ref[2] - tx) ** 2 + (ref[3] * rx + ref[4] * ry + ref[5] - ty) ** 2)
    res /= len(inl)
    if res > 5.0:
        return None
    return ref, n, bi, res


def test_oracle_ransac_and_fits_against_python_restatement(orc):
    rng = np.random.default_rng(5)
    ref = rng.uniform(50, 950, (40, 2))
    th = math.radians(2.0)
    A = np.array([[1.01 * math.cos(th), -math.sin(th)], [math.sin(th), 0.99 * math.cos(th)]])
    tgt = ref @ A.T + np.array([20.0, -11.5]) + rng.normal(0, 0.1, ref.shape)
    tgt[:9] = rng.uniform(50, 950, (9, 2))                           # outliers
    m = np.hstack([ref, tgt])
    ms = [tuple(float(v) for v in row) for row in m]
    fa, fr = orc.fit_affine(m[10:]), orc.fit_rigid(m[10:])
    assert fa.as_tuple() == py_fit_affine(ms[10:]) and fr.as_tuple() == py_fit_rigid(ms[10:])
    for deg in (np.tile(m[:1], (5, 1)), np.array([[0, 0, 1, 1], [1, 1, 2, 2], [2, 2, 3, 3]], np.float64)):  # collinear
        a, b = orc.fit_affine(deg), py_fit_affine([tuple(map(float, r)) for r in deg])
        assert (a is None) == (b is None) and (a is None or a.as_tuple() == b)
    assert orc.fit_affine(np.array([[0, 0, 1, 1], [1, 1, 2, 2], [2, 2, 3, 3]], np.float64)) is None   # det < 1e-12
    for method in ("affine", "rigid"):
        for threads in (1, 8):
            got, want = orc.ransac(m, method, threads), py_ransac(ms, method, threads)
            assert (got is None) == (want is None)
            if got is not None:
                assert got.transform.as_tuple() == want[0]
                assert (got.matched_stars, got.inliers, got.residual_px) == want[1:]
    junk = np.hstack([rng.uniform(0, 1000, (30, 2)), rng.uniform(0, 1000, (30, 2))])
    assert (orc.ransac(junk, "affine", 8) is None) == (py_ransac([tuple(map(float, r)) for r in junk], "affine", 8) is None)
