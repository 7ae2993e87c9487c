"""Host-side mirror of the reference's FITS data-unit codec over the C ABI
(src-tauri/src/infra/fits/reader.rs:37-101 decode_pixels, writer.rs:82-136).  Header parsing and
file I/O are not part of the path.  No CPU fallback."""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import AstroBurstError, default_context, lib

_BPP = {8: 1, 16: 2, 32: 4, -32: 4, -64: 8}


def decode_pixels(data, bitpix: int, bscale: float = 1.0, bzero: float = 0.0, ctx=None) -> np.ndarray:
    """reader.rs:42-101.  `data`: bytes-like big-endian data unit; trailing bytes that do not fill
    a pixel are ignored (par_chunks_exact).  Unknown BITPIX returns an empty array like the reference."""
    ctx = ctx or default_context()
    if bitpix not in _BPP:
        return np.empty(0, np.float32)
    raw = np.frombuffer(data, dtype=np.uint8)
    n = raw.size // _BPP[bitpix]
    out = np.empty(n, np.float32)
    if n:
        raw = np.ascontiguousarray(raw[:n * _BPP[bitpix]])
        ctx.check(lib().ab_fits_decode(ctx.handle, raw.ctypes.data, n, bitpix, float(bscale),
                                       float(bzero), out.ctypes.data))
    return out


def encode_pixels(data: np.ndarray, bitpix: int = -32, bzero: float = 0.0, bscale: float = 1.0,
                  ctx=None) -> bytes:
    """writer.rs:82-136: f32 pixels -> big-endian BITPIX -32 / 16 / -64 data unit (no padding)."""
    ctx = ctx or default_context()
    if bitpix not in (16, -32, -64):
        raise AstroBurstError(_lib.AB_ERR_UNSUPPORTED, f"BITPIX {bitpix} (writer: 16, -32, -64)")
    a = np.ascontiguousarray(data, dtype=np.float32).ravel()
    out = np.empty(a.size * _BPP[bitpix], np.uint8)
    if a.size:
        ctx.check(lib().ab_fits_encode(ctx.handle, a.ctypes.data, a.size, bitpix, float(bzero),
                                       float(bscale), out.ctypes.data))
    return out.tobytes()
