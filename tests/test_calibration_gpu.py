"""Calibration masters and frame calibration on the GPU vs the oracle, plus the
reference's own unit tests (calibration.rs:344-410) through the CUDA path.  Bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cal():
    from astroburst_b200 import calibration
    return calibration


def _frames(n, rows, cols, seed, nan_frac=0.0, level=1000.0):
    rng = np.random.default_rng(seed)
    fr = rng.normal(level, 15.0, (n, rows, cols)).astype(np.float32)
    if nan_frac:
        m = rng.random(fr.shape) < nan_frac
        fr[m] = np.nan
        fr[rng.random(fr.shape) < nan_frac / 4] = np.inf
    return list(fr)


def test_reference_unit_vectors(cal):
    img = np.array([[110, 120], [130, 140]], np.float32)
    bias = np.full((2, 2), 10, np.float32)
    r = cal.subtract_bias(img, bias)                       # test_subtract_bias
    assert abs(r[0, 0] - 100.0) < 1e-6 and abs(r[1, 1] - 130.0) < 1e-6
    r = cal.subtract_dark(np.full((2, 2), 200, np.float32), np.full((2, 2), 20, np.float32), 2.0)
    assert abs(r[0, 0] - 160.0) < 1e-6                     # test_subtract_dark_with_ratio
    img = np.array([[100, 200], [300, 400]], np.float32)
    r = cal.divide_flat(img, np.array([[0.5, 1.0], [1.5, 2.0]], np.float32))
    assert np.all(np.abs(r - 200.0) < 1e-4)                # test_divide_flat
    r = cal.divide_flat(img, np.array([[0.0, 1.0], [np.nan, 2.0]], np.float32))
    assert abs(r[0, 0] - 100.0) < 1e-4 and np.isfinite(r[0, 1]) and abs(r[1, 0] - 300.0) < 1e-4
    raw = np.arange(110, 200, 10, dtype=np.float32).reshape(3, 3)
    cfg = cal.CalibrationConfig(np.full((3, 3), 10, np.float32), np.full((3, 3), 5, np.float32),
                                np.ones((3, 3), np.float32), 1.0)
    r = cal.calibrate_image(raw, cfg)                      # test_full_calibration_pipeline
    assert abs(r[0, 0] - 95.0) < 1e-4 and abs(r[2, 2] - 175.0) < 1e-4


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 16, 21, 31, 32, 33, 50, 64, 100, 256])
def test_median_combine_matches_oracle(cal, orc, n):
    fr = _frames(n, 24, 40, seed=n, nan_frac=0.05)
    fr[0][0, :] = np.nan          # pixels that are non-finite in one frame
    for f in fr:
        f[1, :5] = np.nan         # pixels that are non-finite in every frame -> 0
    got = cal.median_combine(fr)
    want = orc.median_combine(fr)
    assert np.array_equal(got, want)
    assert np.all(got[1, :5] == 0.0)


def test_master_dark_flat_pipeline_matches_oracle(cal, orc):
    rng = np.random.default_rng(7)
    rows, cols = 64, 96
    bias_fr = _frames(9, rows, cols, 1, level=100.0)
    dark_fr = _frames(12, rows, cols, 2, level=130.0)
    flat_fr = [f * rng.uniform(0.7, 1.3, (rows, cols)).astype(np.float32)
               for f in _frames(15, rows, cols, 3, nan_frac=0.01, level=20000.0)]
    mb = cal.create_master_bias(bias_fr)
    assert np.array_equal(mb, orc.median_combine(bias_fr))
    md = cal.create_master_dark(dark_fr, mb)
    assert np.array_equal(md, orc.median_combine(dark_fr, bias=mb))
    mf = cal.create_master_flat(flat_fr, mb, md)
    want = orc.normalize_flat(orc.median_combine(flat_fr, bias=mb, dark=md, dark_ratio=1.0))
    assert np.array_equal(mf, want)
    assert abs(float(mf[np.isfinite(mf)].mean()) - 1.0) < 1e-3
    raw = _frames(1, rows, cols, 4, nan_frac=0.01, level=1500.0)[0]
    mf2 = mf.copy()
    mf2[0, :4] = [0.0, 5e-5, np.nan, np.inf]   # unusable flat pixels: left undivided
    cfg = cal.CalibrationConfig(mb, md, mf2, 2.5)
    got = cal.calibrate_image(raw, cfg)
    want = orc.calibrate_image(raw, mb, md, mf2, 2.5, True)
    assert np.array_equal(got, want, equal_nan=True)
    assert np.array_equal(cal.subtract_dark(raw, md, 0.5), orc.calibrate_image(raw, None, md, None, 0.5, False),
                          equal_nan=True)


def test_normalize_flat_edge_cases(cal, orc):
    z = np.zeros((8, 8), np.float32)
    z[0, 0] = np.nan
    assert np.array_equal(cal.normalize_flat(z), orc.normalize_flat(z), equal_nan=True)  # untouched
    a = np.full((8, 8), -3.0, np.float32)
    a[2, 2] = 4.0
    assert np.array_equal(cal.normalize_flat(a), orc.normalize_flat(a))


def test_error_behaviour(cal):
    from astroburst_b200 import AstroBurstError
    with pytest.raises(AstroBurstError):
        cal.create_master_bias([])
    with pytest.raises(AstroBurstError):
        cal.create_master_bias([np.zeros((4, 4), np.float32), np.zeros((4, 5), np.float32)])


def test_device_band_entry_point(cal, orc):
    """ab_median_combine_dev on a row band (the multi-GPU sharding unit)."""
    import ctypes as C
    import torch
    import astroburst_b200 as ab
    from astroburst_b200._lib import lib
    ctx = ab.default_context(0)
    fr = _frames(20, 32, 64, seed=11, nan_frac=0.02)
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(f[8:24].copy()).to(dev) for f in fr]
    out = torch.empty((16, 64), dtype=torch.float32, device=dev)
    ptrs = (C.c_void_p * len(t))(*[x.data_ptr() for x in t])
    ctx.check(lib().ab_median_combine_dev(ctx.handle, ptrs, len(t), out.numel(), None, None,
                                          C.c_float(1.0), out.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), orc.median_combine(fr)[8:24])


# ---- sigma_clipped_mean_stack of the batch pipeline (calibration_pipeline.rs:317-378)
@pytest.mark.parametrize("n_frames,rows,cols", [(3, 20, 31), (8, 33, 40), (12, 40, 64), (20, 24, 50), (33, 16, 48),
                                                (64, 12, 40), (100, 9, 33), (200, 6, 21)])
def test_sigma_clipped_mean_stack_matches_oracle(n_frames, rows, cols, orc):
    """Bit-exact image (f32 sum in frame order) and per-frame rejection counts, every frame-count
    class of the kernel (register network 8/16/32/64, rolled sort 128/256)."""
    from astroburst_b200 import calibration

    rng = np.random.default_rng(n_frames)
    fr = rng.normal(1.0, 0.02, (n_frames, rows, cols)).astype(np.float32)
    fr[rng.random(fr.shape) < 0.01] += 3.0                         # outliers
    fr[rng.random(fr.shape) < 0.004] = np.nan                      # NaN takes part (ordered last)
    fr[rng.random(fr.shape) < 0.002] = np.inf
    fr[:, 1, 2] = 0.5                                              # MAD 0 -> sigma < 1e-10 ends the loop
    fr[: n_frames // 2 + 1, 2, 3] = np.nan                         # NaN median -> everything rejected -> 0
    for sl, sh, it in ((2.5, 3.0, 5), (1.0, 1.5, 3), (3.0, 3.0, 1), (2.5, 3.0, 0)):
        want, want_cnt = orc.sigma_clipped_mean_stack(list(fr), sl, sh, it)
        got, got_cnt = calibration.sigma_clipped_mean_stack(list(fr), sl, sh, it)
        assert got_cnt == want_cnt, (sl, sh, it)
        assert np.array_equal(got, want, equal_nan=True), (sl, sh, it)
    assert got[2, 3] == 0.0 or it == 0


def test_sigma_clipped_mean_stack_errors():
    import astroburst_b200 as ab
    from astroburst_b200 import calibration

    with pytest.raises(ab.AstroBurstError):
        calibration.sigma_clipped_mean_stack([])
    with pytest.raises(ab.AstroBurstError):
        calibration.sigma_clipped_mean_stack([np.zeros((4, 4), np.float32), np.zeros((4, 5), np.float32)])
