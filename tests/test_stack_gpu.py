"""GPU parity tests for the sigma-clip stack: CUDA path (through the C ABI) vs the CPU oracle.

Bar: `rejected_pixels` bit-exact; pixels bit-exact on these inputs (the only order-dependent
step, the f64 sum of f32 samples, is exact for any dynamic range < 2^29) and in any case within
1e-5 relative (BASELINE.json north_star tolerance).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # north_star: f32 pixel values within 1e-5 relative


@pytest.fixture(scope="module")
def ab():
    import astroburst_b200 as ab

    ab.default_context()
    return ab


def _assert_pixels(got, want):
    assert got.shape == want.shape
    bad = ~(np.isclose(got, want, rtol=REL_TOL, atol=0.0) | (got == want))
    assert not bad.any(), f"{bad.sum()} pixels outside {REL_TOL} relative"
    return int((got != want).sum())


# ---- the reference's unit tests (combine.rs:199-284) through the CUDA path
def test_reference_unit_vectors(ab, orc):
    cases = [
        ([10.0, 10.1, 9.9, 10.0, 10.2], 3.0, 3.0, 5),
        ([10.0, 10.1, 9.9, 10.0, 500.0], 3.0, 3.0, 5),
        ([100.0, 100.2, 99.8, 100.1, 100.0, 5000.0, 99.9], 2.0, 2.0, 5),
        ([], 3.0, 3.0, 5),
        ([42.0], 3.0, 3.0, 5),
    ]
    for vals, sl, sh, it in cases:
        got = ab.sigma_clip_combine(vals, sl, sh, it)
        want = orc.sigma_clip_combine(vals, sl, sh, it)
        assert got == want, (vals, got, want)
    mean, rej = ab.sigma_clip_combine([10.0, 10.1, 9.9, 10.0, 10.2], 3.0, 3.0, 5)
    assert abs(mean - 10.04) < 0.1 and rej == 0
    mean, rej = ab.sigma_clip_combine([10.0, 10.1, 9.9, 10.0, 500.0], 3.0, 3.0, 5)
    assert mean < 15.0 and rej > 0


def test_stack_identical_and_outlier(ab):
    img = (np.arange(16, dtype=np.float32) * 10.0).reshape(4, 4)
    res = ab.stack_images([img, img, img], ab.StackConfig(align=False))
    assert res.frame_count == 3 and abs(res.image[0, 0]) < 1e-4 and abs(res.image[1, 1] - 50.0) < 1e-4
    clean = np.full((4, 4), 100.0, np.float32)
    noisy = clean.copy()
    noisy[2, 2] = 50000.0
    res = ab.stack_images([clean, clean, clean, noisy, clean], ab.StackConfig(3.0, 3.0, 5, False))
    assert abs(res.image[2, 2] - 100.0) < 1.0 and res.rejected_pixels > 0


def test_error_behaviour(ab):
    with pytest.raises(ab.AstroBurstError):  # bail!("No images to stack")
        ab.stack_images([], ab.StackConfig(align=False))
    # align=true on device row bands cannot be phase-correlated: the _dev entry point refuses it
    import ctypes as C
    from astroburst_b200._lib import lib
    ctx = ab.default_context(0)
    cfg = ab.StackConfig().to_c()  # align defaults to true (types/stacking.rs:13-20)
    ptrs = (C.c_void_p * 1)(1)
    assert lib().ab_stack_sigma_clip_dev(ctx.handle, ptrs, 1, 16, C.byref(cfg), 1, 1, None) != 0
    # a single frame with align=true is just that frame (combine.rs:118-121)
    one = np.arange(16, dtype=np.float32).reshape(4, 4)
    res = ab.stack_images([one], ab.StackConfig())
    assert np.array_equal(res.image, one) and res.offsets == [(0, 0)]


# ---- random single-pixel vectors, every length class (register 8/16/32/64 and deep path)
@pytest.mark.parametrize("n", [2, 3, 4, 5, 7, 8, 9, 10, 15, 16, 17, 20, 31, 32, 33, 48, 63, 64, 65, 100, 128, 200, 256])
def test_single_pixel_random(ab, orc, n):
    rng = np.random.default_rng(n)
    for trial in range(6):
        v = rng.normal(1000.0, 12.0, n).astype(np.float32)
        if trial % 2:
            v[rng.integers(0, n)] += 20000.0
        if trial % 3 == 2 and n > 2:
            v[rng.integers(0, n, max(n // 8, 1))] -= 300.0
        sl, sh = [(3.0, 3.0), (2.0, 2.5), (1.0, 4.0)][trial % 3]
        for it in (0, 1, 2, 5, 8):
            assert ab.sigma_clip_combine(v, sl, sh, it) == orc.sigma_clip_combine(v, sl, sh, it), (n, trial, it)


def test_single_pixel_edge_cases(ab, orc):
    cases = [
        [5.0, 5.0, 5.0, 5.0],                 # MAD = 0 -> sigma floor 1e-10
        [1.0, 1.0, 1.0, 1.0, 2.0],            # everything but the mode rejected
        [0.0, 0.0],
        [-3.0, 7.0],
        [1.0, 2.0, 1000.0, 1001.0],           # even n: upper median (combine.rs:38-40)
        list(np.linspace(-50, 50, 32)),
        [3.4e38, 3.4e38, -3.4e38, 1.0],       # dev overflows to inf
    ]
    for v in cases:
        for it in (0, 1, 5):
            for sl, sh in ((3.0, 3.0), (0.5, 0.5), (0.0, 0.0)):
                assert ab.sigma_clip_combine(v, sl, sh, it) == orc.sigma_clip_combine(v, sl, sh, it), (v, it, sl)
    # Dynamic range > 2^29: the f64 sum of the survivors is no longer exact, so the result
    # depends on summation order -- which in the reference is the unspecified permutation left
    # by select_nth_unstable (combine.rs:39,51).  The CUDA path sums in ascending order.  Once the
    # clip has removed the huge samples the sum is exact again and must match bit for bit.
    v = [1e30, -1e30, 1.0, 2.0, 3.0]
    for it in (1, 5):
        assert ab.sigma_clip_combine(v, 3.0, 3.0, it) == orc.sigma_clip_combine(v, 3.0, 3.0, it) == (2.0, 2)


# ---- whole stacks vs oracle
@pytest.mark.parametrize("n_frames,rows,cols", [(32, 96, 160), (10, 64, 200), (3, 33, 47), (16, 50, 70),
                                                (20, 40, 90), (33, 40, 64), (64, 32, 48), (100, 24, 40),
                                                (256, 16, 24), (1, 8, 8), (2, 8, 8)])
def test_stack_matches_oracle(ab, orc, n_frames, rows, cols):
    from astroburst_b200 import synth

    fr = synth.stack_frames(n_frames, rows, cols, seed=n_frames, nan_border=4, cosmic_fraction=2e-3)
    want, want_rej = orc.stack_images(list(fr), 3.0, 3.0, 5)
    res = ab.stack_images(list(fr), ab.StackConfig(3.0, 3.0, 5, False))
    assert res.rejected_pixels == want_rej
    ndiff = _assert_pixels(res.image, want)
    assert ndiff == 0, f"{ndiff} pixels not bit-identical"


def test_stack_all_nan_pixels_and_ragged(ab, orc):
    from astroburst_b200 import synth

    fr = synth.stack_frames(12, 40, 40, seed=3, nan_border=0)
    fr[:, 5, 5] = np.nan           # all-NaN pixel -> 0.0, no rejections (G3)
    fr[:11, 6, 6] = np.inf         # a single finite sample
    fr[::2, 7, :] = -np.inf
    want, want_rej = orc.stack_images(list(fr), 2.5, 3.5, 4)
    res = ab.stack_images(list(fr), ab.StackConfig(2.5, 3.5, 4, False))
    assert res.image[5, 5] == 0.0 and res.rejected_pixels == want_rej
    assert np.array_equal(res.image, want)
    # ragged frames are cropped to the common minimum (combine.rs:104-113)
    ragged = [fr[0], fr[1][:38, :], fr[2][:, :36]]
    want, want_rej = orc.stack_images([f[:38, :36] for f in ragged], 3.0, 3.0, 5)
    res = ab.stack_images(ragged, ab.StackConfig(3.0, 3.0, 5, False))
    assert res.image.shape == (38, 36) and np.array_equal(res.image, want) and res.rejected_pixels == want_rej


def test_stack_max_iter_and_sigma_sweep(ab, orc):
    from astroburst_b200 import synth

    fr = list(synth.stack_frames(32, 48, 64, seed=11, nan_border=0, cosmic_fraction=5e-3))
    for it in (0, 1, 2, 3, 10):
        for sl, sh in ((3.0, 3.0), (1.5, 2.0)):
            want, want_rej = orc.stack_images(fr, sl, sh, it)
            res = ab.stack_images(fr, ab.StackConfig(sl, sh, it, False))
            assert res.rejected_pixels == want_rej and np.array_equal(res.image, want), (it, sl, sh)


# ---- size-independent properties at the BASELINE size (32 x 4096^2), device-resident
def test_full_size_properties(ab, orc):
    import torch
    from astroburst_b200 import synth

    dev = torch.device("cuda:0")
    n, rows, cols = 32, 4096, 4096
    frames = synth.stack_frames_torch(n, rows, cols, dev)
    cfg = ab.StackConfig(3.0, 3.0, 5, False)
    ctx = ab.default_context(0)
    out = torch.empty((rows, cols), dtype=torch.float32, device=dev)
    rej = torch.zeros(1, dtype=torch.int64, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    ab.stack_images_device([f.data_ptr() for f in frames], rows * cols, cfg, out.data_ptr(), rej.data_ptr(), s)
    torch.cuda.synchronize()
    # (1) frame-order invariance: the result is a function of the sample multiset
    perm = torch.randperm(n).tolist()
    out2 = torch.empty_like(out)
    rej2 = torch.zeros_like(rej)
    ab.stack_images_device([frames[i].data_ptr() for i in perm], rows * cols, cfg, out2.data_ptr(), rej2.data_ptr(), s)
    torch.cuda.synchronize()
    assert torch.equal(out, out2) and int(rej) == int(rej2)
    # (2) bounded by the finite sample range; finite everywhere
    st = torch.stack(frames)
    finite = torch.isfinite(st)
    lo = torch.where(finite, st, torch.full_like(st, float("inf"))).amin(0)
    hi = torch.where(finite, st, torch.full_like(st, float("-inf"))).amax(0)
    assert bool(torch.isfinite(out).all()) and bool((out >= lo).all()) and bool((out <= hi).all())
    # (3) oracle on sampled row bands of the same input (bit-exact), incl. the NaN border
    for r0 in (0, 60, 2048, 4096 - 16):
        band = [f[r0:r0 + 16].cpu().numpy() for f in frames]
        want, _ = orc.stack_images(band, 3.0, 3.0, 5)
        assert np.array_equal(out[r0:r0 + 16].cpu().numpy(), want)
    # (4) host-buffer entry point == device entry point
    host = [f[:256].cpu().numpy() for f in frames]
    res = ab.stack_images(host, cfg)
    assert np.array_equal(res.image, out[:256].cpu().numpy())
    # (5) stacking n copies of one frame returns it (finite pixels) with no rejections
    same = [frames[1]] * n
    ab.stack_images_device([f.data_ptr() for f in same], rows * cols, cfg, out2.data_ptr(), rej2.data_ptr(), s)
    torch.cuda.synchronize()
    assert torch.equal(out2, frames[1]) and int(rej2) == 0


# ---- deep path (33..256 frames, stack_deep.cu): far outliers, ragged / unaligned input, gaps
@pytest.mark.parametrize("n", [33, 100, 256])
def test_deep_far_outliers(ab, orc, n):
    """a far finite outlier must not wreck the f64 moments of the later iterations
    (they are accumulated after iteration 0's clip and never contain it)."""
    rng = np.random.default_rng(7 + n)
    cases = [
        (rng.normal(1000.0, 30.0, n), 1e12),
        (rng.normal(0.05, 0.005, n), 1e7),
        (rng.normal(999.0, 5.0, n), 1e30),
        (rng.normal(100.0, 1.0, n), float(np.finfo(np.float32).max)),
        (rng.normal(1e-3, 1e-4, n), 1e7),
    ]
    for v, outlier in cases:
        v = v.astype(np.float32)
        v[n // 3] = outlier
        for it in (1, 2, 5):
            want = orc.sigma_clip_combine(v, 3.0, 3.0, it)
            assert ab.sigma_clip_combine(v, 3.0, 3.0, it) == want, (n, outlier, it)
            assert 1 <= want[1] < n // 2


@pytest.mark.parametrize("n_frames,rows,cols", [(256, 9, 37), (200, 5, 70), (129, 7, 33), (65, 6, 50), (40, 11, 29)])
def test_deep_ragged_unaligned_and_gaps(ab, orc, n_frames, rows, cols):
    """Pixel counts that are not a multiple of the 32-pixel tile, frames at 4-byte (not 16-byte)
    aligned device addresses, per-pixel varying finite counts (NaN / +-inf gaps), all-NaN pixels."""
    import torch
    from astroburst_b200 import synth

    fr = synth.stack_frames(n_frames, rows, cols, seed=n_frames, nan_border=0, cosmic_fraction=4e-3)
    rng = np.random.default_rng(n_frames)
    gaps = rng.random(fr.shape) < 0.07
    fr[gaps] = np.nan
    fr[rng.random(fr.shape) < 0.002] = np.inf
    fr[rng.random(fr.shape) < 0.002] = -np.inf
    fr[:, 1, 2] = np.nan                     # no finite sample at all -> 0
    fr[: n_frames - 1, 2, 3] = np.nan        # exactly one finite sample
    fr[: n_frames - 2, 3, 4] = np.nan        # two
    fr[::2, 4, :] = np.nan                   # half the frames missing along a row
    want, want_rej = orc.stack_images(list(fr), 3.0, 3.0, 5)
    res = ab.stack_images(list(fr), ab.StackConfig(3.0, 3.0, 5, False))
    assert res.rejected_pixels == want_rej and np.array_equal(res.image, want)
    # device entry point, every frame starting 4 bytes off a 16-byte boundary
    dev = torch.device("cuda:0")
    npix = rows * cols
    pad = (npix + 1 + 3) // 4 * 4
    flat = torch.zeros(n_frames * pad + 4, dtype=torch.float32, device=dev)
    ptrs = []
    for f in range(n_frames):
        v = flat[f * pad + 1: f * pad + 1 + npix]
        v.copy_(torch.from_numpy(fr[f].reshape(-1)))
        assert v.data_ptr() % 16 == 4
        ptrs.append(v.data_ptr())
    out = torch.empty(npix, dtype=torch.float32, device=dev)
    rej = torch.zeros(1, dtype=torch.int64, device=dev)
    ab.stack_images_device(ptrs, npix, ab.StackConfig(3.0, 3.0, 5, False), out.data_ptr(), rej.data_ptr(),
                           torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert int(rej) == want_rej and np.array_equal(out.cpu().numpy().reshape(rows, cols), want)


@pytest.mark.parametrize("n_frames", [33, 35, 37, 39, 41, 47, 49, 50, 55, 57, 63])
def test_frame_counts_between_register_kernels(ab, orc, n_frames):
    """33..63 frames without a kernel of their own run the next register kernel up, the missing
    frames delivered as NaN by the tensor map: same image and reject count as the oracle, on a
    constant-stride cube (the padded route), with a ragged tail, a padded plane stride, per-pixel
    gaps beyond the 8 the specialised tails cover, and all-NaN pixels."""
    import torch
    from astroburst_b200 import synth

    rows, cols = 21, 101                    # 2,121 px = 66 tiles + 9 px
    fr = synth.stack_frames(n_frames, rows, cols, seed=1000 + n_frames, nan_border=2, cosmic_fraction=3e-3)
    rng = np.random.default_rng(n_frames)
    fr[rng.random(fr.shape) < 0.03] = np.nan
    fr[rng.random(fr.shape) < 0.001] = np.inf
    fr[:, 5, 6] = np.nan
    fr[: n_frames - 1, 6, 7] = np.nan
    fr[::2, 7, :] = np.nan                  # half the frames missing along a row
    fr[: n_frames - 20, 8, :50] = np.nan    # many missing
    want, want_rej = orc.stack_images(list(fr), 3.0, 3.0, 5)
    dev = torch.device("cuda:0")
    npix = rows * cols
    cfg = ab.StackConfig(3.0, 3.0, 5, False)
    stream = torch.cuda.current_stream().cuda_stream
    for stride in (npix + (-npix) % 4, npix + 4 * 37):       # tight (16-byte aligned planes) and padded
        cube = torch.full((n_frames, stride), 7.0, dtype=torch.float32, device=dev)
        cube[:, :npix] = torch.from_numpy(fr.reshape(n_frames, -1))
        if (stride * 4) % 16:
            continue
        ptrs = [cube[f].data_ptr() for f in range(n_frames)]
        assert all(p % 16 == 0 for p in ptrs)
        out = torch.empty(npix, dtype=torch.float32, device=dev)
        rej = torch.full((1,), 12345, dtype=torch.int64, device=dev)   # stale value: the launch must overwrite it
        ab.stack_images_device(ptrs, npix, cfg, out.data_ptr(), rej.data_ptr(), stream)
        torch.cuda.synchronize()
        assert int(rej) == want_rej and np.array_equal(out.cpu().numpy().reshape(rows, cols), want), stride
    res = ab.stack_images(list(fr), cfg)    # host buffers
    assert res.rejected_pixels == want_rej and np.array_equal(res.image, want)


def test_deep_band_properties_at_config5_size(ab, orc):
    """256 frames x (64 x 4096) band (1/64 of BASELINE config 5): oracle on sampled rows,
    frame-order invariance, sigma / iteration sweep."""
    import torch

    dev = torch.device("cuda:0")
    n, rows, cols = 256, 64, 4096
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    cube = torch.randn((n, rows, cols), generator=g, device=dev) * 12.0 + 1000.0
    idx = torch.randint(0, cube.numel(), (cube.numel() // 5000,), generator=g, device=dev)
    cube.view(-1)[idx] += torch.rand(idx.numel(), generator=g, device=dev) * 45000.0 + 5000.0
    cube[::8, :3, :] = float("nan")          # frame-uniform gap rows (registration borders)
    s = torch.cuda.current_stream().cuda_stream
    out = torch.empty((rows, cols), dtype=torch.float32, device=dev)
    rej = torch.zeros(1, dtype=torch.int64, device=dev)
    for (sl, sh, it) in ((3.0, 3.0, 5), (2.0, 2.5, 2), (3.0, 3.0, 0)):
        cfg = ab.StackConfig(sl, sh, it, False)
        ab.stack_images_device([cube[f].data_ptr() for f in range(n)], rows * cols, cfg, out.data_ptr(),
                               rej.data_ptr(), s)
        torch.cuda.synchronize()
        for r0 in (0, 2, 40):
            band = [cube[f, r0:r0 + 2].cpu().numpy() for f in range(n)]
            want, _ = orc.stack_images(band, sl, sh, it)
            assert np.array_equal(out[r0:r0 + 2].cpu().numpy(), want), (sl, sh, it, r0)
        want_all, want_rej = orc.stack_images([cube[f, :8].cpu().numpy() for f in range(n)], sl, sh, it)
        rej8 = torch.zeros(1, dtype=torch.int64, device=dev)
        out8 = torch.empty((8, cols), dtype=torch.float32, device=dev)
        ab.stack_images_device([cube[f].data_ptr() for f in range(n)], 8 * cols, cfg, out8.data_ptr(),
                               rej8.data_ptr(), s)
        torch.cuda.synchronize()
        assert int(rej8) == want_rej and np.array_equal(out8.cpu().numpy(), want_all)
    perm = torch.randperm(n).tolist()
    out2 = torch.empty_like(out)
    rej2 = torch.zeros_like(rej)
    ab.stack_images_device([cube[f].data_ptr() for f in range(n)], rows * cols, ab.StackConfig(3.0, 3.0, 5, False),
                           out.data_ptr(), rej.data_ptr(), s)
    ab.stack_images_device([cube[f].data_ptr() for f in perm], rows * cols, ab.StackConfig(3.0, 3.0, 5, False),
                           out2.data_ptr(), rej2.data_ptr(), s)
    torch.cuda.synchronize()
    assert torch.equal(out, out2) and int(rej) == int(rej2)
