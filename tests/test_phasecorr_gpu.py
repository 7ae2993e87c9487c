"""Phase correlation + aligned stack on the GPU vs the oracle.

Parity for this row is tolerance-based: the reference's FFT is rustfft (un-vendored); cuFFT and the
oracle's radix-2 DFT agree with it to f64 rounding (shifts ~1e-12 px).  Tolerances used here:
1e-6 px on dx/dy, 1e-6 relative on the confidence; area_downsample is bit-exact.

Sign convention: the reference's code (complex.rs:27-33: a * conj(b); fft.rs forward = rustfft
plan_fft_forward) returns the NEGATIVE of the content displacement -- restated literally by the
oracle and the CUDA path.  The reference's own tests (phase_correlation.rs:204-218,
align.rs:187-224) assert the opposite sign on a near-periodic pattern whose phase-only correlation
does not even recover the magnitude; they are ported here on a broadband star field, asserting what
the literal code returns."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def al():
    from astroburst_b200 import alignment
    return alignment


def star_field(rows, cols, seed, n_stars=None):
    rng = np.random.default_rng(seed)
    img = rng.normal(1000.0, 10.0, (rows, cols)).astype(np.float32)
    for _ in range(n_stars or rows * cols // 400):
        y, x = rng.integers(3, rows - 3), rng.integers(3, cols - 3)
        img[y - 1:y + 2, x - 1:x + 2] += np.float32(rng.uniform(500, 5000))
    return img


def shift_array(img, dy, dx):  # phase_correlation.rs:176-194
    out = np.zeros_like(img)
    R, C = img.shape
    out[max(dy, 0):R + min(dy, 0), max(dx, 0):C + min(dx, 0)] = \
        img[max(-dy, 0):R + min(-dy, 0), max(-dx, 0):C + min(-dx, 0)]
    return out


def close(got, want):
    gd, wd = (got.dx, got.dy, got.confidence), want
    assert abs(gd[0] - wd[0]) < 1e-6 and abs(gd[1] - wd[1]) < 1e-6, (gd, wd)
    assert abs(gd[2] - wd[2]) <= 1e-6 * max(1.0, abs(wd[2])), (gd, wd)


@pytest.mark.parametrize("shape,out", [((64, 64), (64, 64)), ((4, 4), (2, 2)), ((1000, 1300), (512, 512)),
                                       ((700, 333), (512, 333)), ((513, 2049), (100, 37))])
def test_area_downsample_bit_exact(al, orc, shape, out):
    rng = np.random.default_rng(1)
    img = rng.normal(100.0, 30.0, shape).astype(np.float32)
    img[rng.random(shape) < 0.01] = np.nan
    img[:3, :5] = np.inf
    got = al.area_downsample(img, *out)
    assert np.array_equal(got, orc.area_downsample(img, *out), equal_nan=True)
    if shape == (4, 4):  # downsample.rs:60-67 test_halve
        ramp = np.arange(16, dtype=np.float32).reshape(4, 4)
        assert abs(al.area_downsample(ramp, 2, 2)[0, 0] - 2.5) < 1e-4


@pytest.mark.parametrize("shape,shift", [((128, 128), (0, 0)), ((256, 256), (10, -5)), ((300, 280), (7, -4)),
                                         ((512, 512), (-31, 12)), ((1200, 1500), (40, -25)),
                                         ((600, 2100), (-9, 60)), ((2048, 2048), (3, 2))])
def test_phase_correlate_matches_oracle(al, orc, shape, shift):
    ref = star_field(*shape, seed=shape[0] + shift[0])
    tgt = shift_array(ref, *shift)
    got = al.phase_correlate(ref, tgt)
    close(got, orc.phase_correlate(ref, tgt))
    assert got.confidence > 2.0
    if max(shape) <= 512:
        # single correlation: the literal reference code returns minus the content displacement
        assert abs(got.dy + shift[0]) < 0.5 and abs(got.dx + shift[1]) < 0.5, (got, shift)
    # (coarse-to-fine cases: the refinement crop is centred with the same negated coarse estimate,
    # phase_correlation.rs:68-73, so the literal code does not converge to the displacement; the
    # parity check above is what those cases are for)


def test_reference_edge_cases(al, orc):
    img = star_field(128, 128, 3)
    r = al.phase_correlate(img, img)                      # test_identical_images
    assert abs(r.dx) < 0.5 and abs(r.dy) < 0.5
    bad = star_field(64, 64, 4)
    bad[10, 10], bad[20, 30], bad[5, 5] = np.nan, np.inf, -np.inf
    r = al.phase_correlate(bad, bad)                      # test_nan_no_panic
    assert np.isfinite(r.dx) and np.isfinite(r.dy)
    close(r, orc.phase_correlate(bad, bad))
    const = np.full((64, 64), 100.0, np.float32)          # test_constant_image
    r = al.phase_correlate(const, const)
    assert (r.dx, r.dy, r.confidence) == (0.0, 0.0, 0.0)
    # different sizes are cropped to the common minimum (phase_correlation.rs:26-41)
    a, b = star_field(200, 260, 5), star_field(230, 250, 5)
    close(al.phase_correlate(a, b), orc.phase_correlate(a, b))


def test_stack_images_with_align(orc):
    """stack_images(align=true), combine.rs:118-147: offsets identical, pixels to 1e-5 relative
    (the shifts differ from the oracle's by ~1e-12 px, so a few bicubic weights differ in their
    last f64 bit)."""
    import astroburst_b200 as ab
    rows, cols, n = 240, 320, 6
    base = star_field(rows + 40, cols + 40, 9)
    rng = np.random.default_rng(2)
    shifts = [(0, 0), (3, -2), (-5, 4), (1, 1), (7, 0), (-2, -6)]
    frames = [(base[20 + dy:20 + dy + rows, 20 + dx:20 + dx + cols]
               + rng.normal(0, 5, (rows, cols))).astype(np.float32) for dy, dx in shifts]
    res = ab.stack_images(frames, ab.StackConfig(3.0, 3.0, 5, True))
    want, want_rej, want_off = orc.stack_images_aligned(frames, 3.0, 3.0, 5)
    assert res.offsets == want_off
    assert res.offsets == [(dy, dx) for dy, dx in shifts]  # frame(y) = base(y + dy): content moved by -dy
    assert np.allclose(res.image, want, rtol=1e-5, atol=1e-4)
    assert abs(res.rejected_pixels - want_rej) <= max(4, want_rej // 1000)
    # (with the literal sign convention the frames are shifted AWAY from the reference frame --
    # see the module docstring and DESIGN.md; reproduced faithfully, not "fixed")


def test_align_2_registers_frames_onto_the_reference(al):
    """align = 2 (opt-in, outside parity): operands of the phase correlation swapped, so frames are
    shifted ONTO frame 0.  A stack of integer-shifted copies then reproduces frame 0 in the interior,
    while align = 1 (upstream's sign convention, bit parity) smears it."""
    import astroburst_b200 as ab
    import ctypes as C
    from astroburst_b200._lib import StackConfigC, lib

    base = star_field(256, 256, 8)
    frames = [base] + [shift_array(base, dy, dx) for dy, dx in ((3, -2), (-4, 5), (6, 1), (-1, -6))]
    ctx = ab.default_context()

    def run(align):
        cfg = StackConfigC(3.0, 3.0, 5, align)
        ptrs = (C.c_void_p * len(frames))(*[f.ctypes.data for f in frames])
        out = np.empty_like(base)
        rej = C.c_uint64(0)
        offs = np.zeros((len(frames), 2), np.int32)
        ctx.check(lib().ab_stack_images(ctx.handle, ptrs, len(frames), 256, 256, C.byref(cfg), out.ctypes.data,
                                        C.byref(rej), offs.ctypes.data))
        return out, offs

    fixed, offs2 = run(2)
    upstream, offs1 = run(1)
    assert [tuple(o) for o in offs2[1:]] == [(3, -2), (-4, 5), (6, 1), (-1, -6)]   # the content displacement
    assert [tuple(o) for o in offs1[1:]] == [(-3, 2), (4, -5), (-6, -1), (1, 6)]   # upstream: its negative
    core = (slice(16, -16), slice(16, -16))
    err2 = np.abs(fixed[core] - base[core]).max()
    err1 = np.abs(upstream[core] - base[core]).max()
    assert err2 < 1e-2 * base[core].max() and err1 > 10 * max(err2, 1.0)
