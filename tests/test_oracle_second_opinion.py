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
