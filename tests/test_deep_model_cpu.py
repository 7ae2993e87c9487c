"""CPU model of the deep-stack kernel's algorithm (astroburst_b200/csrc/stack_deep.cu) checked
against the oracle (sigma_clip_combine, combine.rs:14-92).

The CUDA kernel cannot run here; this restates what it does per pixel -- four lanes of R registers,
in-lane sort, two bitonic merge levels across the lane quad (flip / straight / half-cleaner
stages), the crossing search for the upper-median MAD, the quad probes of the clip and the f64
moments about the median taken AFTER iteration 0's clip and downdated later -- in numpy scalar code
with the kernel's exact arithmetic (f32 compares, f64 sums), and fuzzes it against the oracle:
identical value and reject count, including the far-outlier cases of ADVICE r1 (1e12, 1e30, FLT_MAX)
that broke the previous kernel's running sums.
"""
import numpy as np
import pytest

f32 = np.float32
INF = f32(np.inf)


def _halfclean(regs):
    """in-lane half-cleaners, strides R/2..1 (ascending), on one lane's registers"""
    R = len(regs)
    stride = R // 2
    while stride >= 1:
        for i in range(R // 2):
            lo = (i // stride) * 2 * stride + (i % stride)
            hi = lo + stride
            a, b = regs[lo], regs[hi]
            regs[lo], regs[hi] = min(a, b), max(a, b)
        stride //= 2


def _flip(lanes, xm, upper_bit):
    R = len(lanes[0])
    old = [list(x) for x in lanes]
    for lane in range(4):
        upper = (lane & upper_bit) != 0
        part = old[lane ^ xm]
        for i in range(R // 2):
            y = part[R - 1 - i]
            x = part[i]
            lanes[lane][i] = max(old[lane][i], y) if upper else min(old[lane][i], y)
            lanes[lane][R - 1 - i] = max(old[lane][R - 1 - i], x) if upper else min(old[lane][R - 1 - i], x)


def _straight(lanes, xm, upper_bit):
    R = len(lanes[0])
    old = [list(x) for x in lanes]
    for lane in range(4):
        upper = (lane & upper_bit) != 0
        for i in range(R):
            x = old[lane ^ xm][i]
            lanes[lane][i] = max(old[lane][i], x) if upper else min(old[lane][i], x)


def quad_sort(samples, R):
    """samples: 4R values (non-finite already +inf), lane l gets [R l, R l + R) -> sorted ranks"""
    lanes = [sorted(samples[R * l:R * l + R]) for l in range(4)]
    for level in range(2):
        _flip(lanes, 3 if level else 1, 2 if level else 1)
        if level:
            _straight(lanes, 1, 1)
        for lane in lanes:
            _halfclean(lane)
    return [v for lane in lanes for v in lane]


def _div_small(x, n):
    return np.float64(x) / np.float64(n)  # the kernel's Markstein division is correctly rounded


def deep_model(values, sigma_low, sigma_high, max_iter, R):
    sigma_low, sigma_high = f32(sigma_low), f32(sigma_high)
    v = np.asarray(values, dtype=f32)
    pad = np.full(4 * R, INF, dtype=f32)
    pad[:len(v)] = np.where(np.isfinite(v), v, INF)
    n = int(np.isfinite(pad).sum())
    S = quad_sort([f32(x) for x in pad], R)
    assert S == sorted(S), "merge network did not sort"
    h = n >> 1
    med = S[h] if n > 0 else f32(0)
    result, rejected = med, 0
    a, b = 0, n
    active = n >= 2 and max_iter > 0
    sigma = f32(0)
    last_center = f32(0)
    with np.errstate(all="ignore"):
        if active:
            W = n - h
            lo_t, hi_t = 0, W
            for _ in range(int(np.log2(R)) + 2):
                mid = (lo_t + hi_t) >> 1
                if lo_t < hi_t:
                    if f32(S[mid + h] - med) >= f32(med - S[mid]):
                        hi_t = mid
                    else:
                        lo_t = mid + 1
            mad = INF
            if lo_t < W:
                mad = max(f32(med - S[lo_t]), f32(S[lo_t + h] - med))
            if lo_t > 0:
                mad = min(mad, max(f32(med - S[lo_t - 1]), f32(S[lo_t - 1 + h] - med)))
            sigma = max(f32(np.float64(mad) * 1.4826), f32(1e-10))
        c0 = np.float64(med)
        D1 = D2 = D2_fresh = np.float64(0)
        center = med
        it = 0
        while True:
            removed = 0
            if it < max_iter and active:
                if b - a < 2:
                    active = False
                if active:
                    last_center = center
                lo = f32(-sigma_low * sigma)
                hi = f32(sigma_high * sigma)
                na, nb = a, b
                search = active
                while search:  # quad probes: four candidates per step
                    cnt = 0
                    for l in range(4):
                        g = na + l
                        if g < nb and not (f32(S[g] - center) >= lo):
                            cnt += 1
                        else:
                            break
                    na += cnt
                    search = cnt == 4 and na < nb
                search = active and na < nb
                while search:
                    cnt = 0
                    for l in range(4):
                        g = nb - 1 - l
                        if g >= na and not (f32(S[g] - center) <= hi):
                            cnt += 1
                        else:
                            break
                    nb -= cnt
                    search = cnt == 4 and nb > na
                a_old, b_old = a, b
                if active:
                    removed = (b - a) - (nb - na)
                    rejected += removed
                    a, b = na, nb
                    if removed == 0:
                        active = False
                if it > 0 and removed > 0:
                    r1 = r2 = np.float64(0)
                    for g in list(range(a_old, a)) + list(range(b, b_old)):
                        d = np.float64(S[g]) - c0
                        r1 += d
                        r2 += d * d
                    D1 -= r1
                    D2 -= r2
            first = it == 0
            stale = (not first) and active and D2 * 1024.0 < D2_fresh
            if first or stale:
                d = np.array([np.float64(S[g]) - c0 for g in range(a, b)], dtype=np.float64)
                D1, D2 = d.sum(), (d * d).sum()
                D2_fresh = D2
            if it + 1 >= max_iter or not active:
                break
            ln = max(b - a, 1)
            mean = _div_small(np.float64(ln) * c0 + D1, ln)
            dm = mean - c0
            ss = max((D2 - 2.0 * dm * D1) + np.float64(ln) * dm * dm, 0.0)
            center = f32(mean)
            sigma = max(f32(np.sqrt(_div_small(ss, ln - 1 if ln > 2 else 1))), f32(1e-10))
            it += 1
        ln = b - a
        if n >= 2:
            if ln == 0:
                result = last_center if np.isfinite(last_center) else f32(0)
            else:
                result = f32(_div_small(np.float64(ln) * c0 + D1, ln))
    return float(result), rejected


def _r_for(n):
    return 8 if n <= 32 else 16 if n <= 64 else 32 if n <= 128 else 64


@pytest.mark.parametrize("n", [2, 5, 31, 33, 40, 64, 65, 100, 128, 129, 200, 255, 256])
def test_model_matches_oracle_random(orc, n):
    rng = np.random.default_rng(1000 + n)
    for trial in range(12):
        v = rng.normal(1000.0, 12.0, n).astype(f32)
        if trial % 2:
            v[rng.integers(0, n, max(1, n // 40))] += f32(20000.0)
        if trial % 3 == 2:
            v[rng.integers(0, n, max(n // 8, 1))] -= f32(300.0)
        if trial % 4 == 3:
            v[rng.integers(0, n, max(n // 10, 1))] = np.nan
        sl, sh = [(3.0, 3.0), (2.0, 2.5), (1.0, 4.0)][trial % 3]
        for it in (0, 1, 2, 5, 8):
            fin = v[np.isfinite(v)]
            want = orc.sigma_clip_combine(fin, sl, sh, it)
            got = deep_model(v, sl, sh, it, _r_for(n))
            assert got == want, (n, trial, it, got, want)


@pytest.mark.parametrize("n", [33, 100, 256])
def test_model_far_outliers(orc, n):
    """ADVICE r1 (stack_deep.cu:381): one far outlier must not wreck the running moments."""
    rng = np.random.default_rng(7 + n)
    cases = [
        (rng.normal(1000.0, 30.0, n).astype(f32), f32(1e12)),
        (rng.normal(0.05, 0.005, n).astype(f32), f32(1e7)),
        (rng.normal(999.0, 5.0, n).astype(f32), f32(1e30)),
        (rng.normal(100.0, 1.0, n).astype(f32), np.finfo(f32).max),
        (rng.normal(1e-3, 1e-4, n).astype(f32), f32(1e7)),
    ]
    for v, out in cases:
        v = v.copy()
        v[n // 3] = out
        for it in (1, 2, 5):
            want = orc.sigma_clip_combine(v, 3.0, 3.0, it)
            got = deep_model(v, 3.0, 3.0, it, _r_for(n))
            assert got == want, (n, float(out), it, got, want)
            assert want[1] >= 1 and want[1] < n // 2  # the oracle rejects the outlier, not everything


def test_model_edge_cases(orc):
    cases = [
        [5.0] * 40,                              # MAD = 0 -> sigma floor
        [1.0] * 39 + [2.0],                      # everything but the mode rejected
        list(np.linspace(-50, 50, 64)),
        [0.0] * 33,
        [1.0, 2.0] * 20 + [1000.0, 1001.0] * 3,
        list(np.r_[np.zeros(20), np.ones(20) * 1e-30, 1.0]),
    ]
    for v in cases:
        for it in (0, 1, 5):
            for sl, sh in ((3.0, 3.0), (0.5, 0.5), (0.0, 0.0)):
                want = orc.sigma_clip_combine(np.asarray(v, f32), sl, sh, it)
                got = deep_model(v, sl, sh, it, _r_for(len(v)))
                assert got == want, (v[:4], len(v), it, sl, got, want)
