"""FITS data-unit decode / encode on the GPU vs the oracle: bit-exact, plus the
reference's unit tests (reader.rs:570-616)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fits():
    from astroburst_b200 import fits
    return fits


def test_reference_unit_vectors(fits):
    p = fits.decode_pixels(bytes([0x01, 0x00, 0xFF, 0xFF]), 16)          # test_decode_pixels_i16
    assert p.size == 2 and abs(p[0] - 256.0) < 1e-6 and abs(p[1] + 1.0) < 1e-6
    p = fits.decode_pixels(bytes([0x3F, 0x80, 0x00, 0x00]), -32)          # test_decode_pixels_f32
    assert p.size == 1 and abs(p[0] - 1.0) < 1e-6
    p = fits.decode_pixels(bytes([100]), 8, 2.0, 10.0)                   # test_decode_pixels_with_scaling
    assert abs(p[0] - 210.0) < 1e-6
    pi = np.float32(np.pi)                                               # ..._identity_f32_fast_path
    assert fits.decode_pixels(np.array([pi]).astype(">f4").tobytes(), -32)[0] == pi
    assert fits.decode_pixels(b"\x00" * 8, 64).size == 0                 # unknown BITPIX -> empty


@pytest.mark.parametrize("bitpix", [8, 16, 32, -32, -64])
@pytest.mark.parametrize("scaling", [(1.0, 0.0), (1.0, 32768.0), (0.0625, -3.5), (1.0 + 1e-16, 0.0)])
def test_decode_matches_oracle(fits, orc, bitpix, scaling):
    rng = np.random.default_rng(abs(bitpix) + int(scaling[1]))
    n = 100003
    nbytes = n * {8: 1, 16: 2, 32: 4, -32: 4, -64: 8}[bitpix]
    raw = rng.integers(0, 256, nbytes, dtype=np.uint8).tobytes()   # every bit pattern incl. NaN / inf
    tail = b"" if bitpix == 8 else b"\x01"                         # a trailing partial pixel is ignored
    got = fits.decode_pixels(raw + tail, bitpix, *scaling)
    want = orc.fits_decode(raw, bitpix, *scaling)
    assert got.size == n
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("bitpix,bzero,bscale", [(-32, 0.0, 1.0), (-64, 0.0, 1.0), (16, 0.0, 1.0),
                                                 (16, 32768.0, 1.0), (16, 1234.5, 0.37)])
def test_encode_matches_oracle_and_round_trips(fits, orc, bitpix, bzero, bscale):
    rng = np.random.default_rng(3)
    a = rng.normal(20000.0, 15000.0, 70001).astype(np.float32)
    a[:6] = [np.nan, np.inf, -np.inf, 0.5, -0.5, 1e9]
    got = fits.encode_pixels(a, bitpix, bzero, bscale)
    assert got == orc.fits_encode(a, bitpix, bzero, bscale)
    back = fits.decode_pixels(got, bitpix, bscale, bzero)
    if bitpix in (-32, -64):
        assert np.array_equal(back.view(np.uint32), a.view(np.uint32))  # lossless
    else:
        fin = np.isfinite(a) & (np.abs((a - bzero) / bscale) < 32000)
        assert np.max(np.abs(back[fin] - a[fin])) <= 0.5 * abs(bscale) + 1e-2


@pytest.mark.parametrize("case", ["i16_bzero", "mixed", "f32_identity", "ragged_many_bands"])
def test_stack_fits_units_equals_decode_then_stack(fits, orc, case, monkeypatch):
    """ab_stack_fits_units (data units over PCIe in the file's sample type, decoded on the device)
    against the oracle's decode_pixels + stack_images: same image, same reject count."""
    import astroburst_b200 as ab
    from astroburst_b200 import synth

    rng = np.random.default_rng(7)
    if case == "ragged_many_bands":
        monkeypatch.setenv("AB_BAND_MB", "16")  # (default 384 MiB: this stack would be one band)
        n, rows, cols = 40, 700, 1031          # several bands, tail pixels, the padded 40-frame kernel
    else:
        n, rows, cols = 12, 61, 83
    fr = synth.stack_frames(n, rows, cols, seed=77, nan_border=0, cosmic_fraction=2e-3)
    units, bitpix, bscale, bzero = [], [], [], []
    for f in range(n):
        kind = {"i16_bzero": 16, "f32_identity": -32, "ragged_many_bands": 16}.get(case) or (8, 16, 32, -32, -64)[f % 5]
        if kind == 16:
            bs, bz = 1.0, 32768.0
            q = np.clip(np.rint((fr[f].astype(np.float64) - bz) / bs), -32768, 32767).astype(">i2")
        elif kind == 8:
            bs, bz = 8.0, 0.0
            q = np.clip(np.rint(fr[f] / bs), 0, 255).astype(np.uint8)
        elif kind == 32:
            bs, bz = 0.001, 10.0
            q = np.rint((fr[f].astype(np.float64) - bz) / bs).astype(">i4")
        elif kind == -32:
            bs, bz = 1.0, 0.0
            q = fr[f].astype(">f4")
            q.reshape(-1)[rng.integers(0, q.size, 5)] = np.nan       # non-finite samples survive the decode
        else:
            bs, bz = 1.0 + 1e-3, -0.5
            q = fr[f].astype(">f8")
        units.append(np.frombuffer(q.tobytes(), np.uint8))
        bitpix.append(kind); bscale.append(bs); bzero.append(bz)
    decoded = [orc.fits_decode(u.tobytes(), b, s, z).reshape(rows, cols) for u, b, s, z in zip(units, bitpix, bscale, bzero)]
    want, want_rej = orc.stack_images(decoded, 3.0, 3.0, 5)
    res = ab.stack_fits_units(units, bitpix, bscale, bzero, rows, cols, ab.StackConfig(3.0, 3.0, 5, False))
    assert res.rejected_pixels == want_rej and np.array_equal(res.image, want, equal_nan=True)
    if case == "ragged_many_bands":         # the f32 host-buffer call over the same 8 bands
        res32 = ab.stack_images(decoded, ab.StackConfig(3.0, 3.0, 5, False))
        assert res32.rejected_pixels == want_rej and np.array_equal(res32.image, want, equal_nan=True)
    with pytest.raises(ab.AstroBurstError):
        ab.stack_fits_units(units, [64] * n, bscale, bzero, rows, cols, ab.StackConfig(3.0, 3.0, 5, False))
    with pytest.raises(ab.AstroBurstError):
        ab.stack_fits_units(units, bitpix, bscale, bzero, rows, cols, ab.StackConfig(3.0, 3.0, 5, True))
