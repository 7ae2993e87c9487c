"""Deterministic synthetic inputs for the tests and bench.py (our own seeded generator; the
reference's core/synth uses rand 0.8.5 StdRng whose stream is not reproducible here).

Stack frames: sky 1000 + Gaussian stars (sigma 1.3 px, power-law flux) +
N(0, 12) noise per frame + 0.02 % cosmic-ray hits per frame + a NaN border strip in frames
k % 8 == 0 (exercises the finite filter, combine.rs:170-175).  Values come from continuous
distributions, so exact ties (where f32 clip decisions could legitimately flip) do not occur.
"""
from __future__ import annotations

import numpy as np


def star_catalogue(rows: int, cols: int, n_stars: int, seed: int = 42):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, cols, n_stars)
    y = rng.uniform(0, rows, n_stars)
    # power-law fluxes between 1e2 and 5e4
    u = rng.uniform(0, 1, n_stars)
    flux = 1e2 * (5e4 / 1e2) ** (u ** 3)
    return x, y, flux


def render_stars(rows: int, cols: int, x, y, flux, sigma: float = 1.3) -> np.ndarray:
    img = np.zeros((rows, cols), np.float64)
    r = int(np.ceil(5 * sigma))
    for sx, sy, f in zip(x, y, flux):
        cx, cy = int(round(sx)), int(round(sy))
        x0, x1 = max(cx - r, 0), min(cx + r + 1, cols)
        y0, y1 = max(cy - r, 0), min(cy + r + 1, rows)
        if x0 >= x1 or y0 >= y1:
            continue
        yy, xx = np.mgrid[y0:y1, x0:x1]
        img[y0:y1, x0:x1] += f / (2 * np.pi * sigma * sigma) * np.exp(
            -((xx - sx) ** 2 + (yy - sy) ** 2) / (2 * sigma * sigma))
    return img


def stack_frames(n_frames: int, rows: int, cols: int, seed: int = 42, nan_border: int = 8,
                 n_stars: int | None = None, cosmic_fraction: float = 2e-4,
                 border: tuple | None = None) -> np.ndarray:
    """-> float32 [n_frames, rows, cols].  `border` = (top, bottom, left, right) NaN strip widths
    in frames k % 8 == 0, instead of the symmetric `nan_border` (row-band samples of a larger
    frame)."""
    if n_stars is None:
        n_stars = max(int(6000 * rows * cols / 4096 ** 2), 4)
    x, y, flux = star_catalogue(rows, cols, n_stars, seed)
    base = 1000.0 + render_stars(rows, cols, x, y, flux)
    out = np.empty((n_frames, rows, cols), np.float32)
    for k in range(n_frames):
        rng = np.random.default_rng(1000 + k + 7919 * seed)
        f = base + rng.normal(0.0, 12.0, (rows, cols))
        n_cr = int(round(cosmic_fraction * rows * cols))
        if n_cr:
            idx = rng.integers(0, rows * cols, n_cr)
            f.reshape(-1)[idx] += rng.uniform(5000.0, 50000.0, n_cr)
        f = f.astype(np.float32)
        if border is not None and k % 8 == 0:
            t, bt, lf, rt = (int(x) for x in border)
            if t:
                f[:t, :] = np.nan
            if bt:
                f[-bt:, :] = np.nan
            if lf:
                f[:, :lf] = np.nan
            if rt:
                f[:, -rt:] = np.nan
        elif nan_border and k % 8 == 0:
            b = min(nan_border, rows // 4, cols // 4)
            if b:
                f[:b, :] = np.nan
                f[-b:, :] = np.nan
                f[:, :b] = np.nan
                f[:, -b:] = np.nan
        out[k] = f
    return out


def stack_frames_torch(n_frames: int, rows: int, cols: int, device, seed: int = 42,
                       nan_border: int = 64, n_stars: int | None = None,
                       cosmic_fraction: float = 2e-4):
    """Same recipe generated on the GPU with torch (bench-size inputs; not bit-identical to
    the numpy generator).  -> list of n_frames float32 [rows, cols] device tensors."""
    import torch

    if n_stars is None:
        n_stars = max(int(6000 * rows * cols / 4096 ** 2), 4)
    x, y, flux = star_catalogue(rows, cols, n_stars, seed)
    base = torch.full((rows, cols), 1000.0, dtype=torch.float32, device=device)
    # stars: scatter small Gaussian stamps (vectorised over stars, 13x13 stamp)
    r = 6
    dy, dx = torch.meshgrid(torch.arange(-r, r + 1, device=device),
                            torch.arange(-r, r + 1, device=device), indexing="ij")
    sx = torch.tensor(x, device=device, dtype=torch.float64)
    sy = torch.tensor(y, device=device, dtype=torch.float64)
    fl = torch.tensor(flux, device=device, dtype=torch.float64)
    cx, cy = sx.round().long(), sy.round().long()
    px = cx[:, None, None] + dx[None]
    py = cy[:, None, None] + dy[None]
    val = fl[:, None, None] / (2 * np.pi * 1.3 * 1.3) * torch.exp(
        -((px - sx[:, None, None]) ** 2 + (py - sy[:, None, None]) ** 2) / (2 * 1.3 * 1.3))
    ok = (px >= 0) & (px < cols) & (py >= 0) & (py < rows)
    base.view(-1).index_add_(0, (py * cols + px)[ok], val[ok].float())
    frames = []
    for k in range(n_frames):
        g = torch.Generator(device=device)
        g.manual_seed(1000 + k + 7919 * seed)
        f = base + torch.randn((rows, cols), generator=g, device=device) * 12.0
        n_cr = int(round(cosmic_fraction * rows * cols))
        if n_cr:
            idx = torch.randint(0, rows * cols, (n_cr,), generator=g, device=device)
            amp = torch.rand((n_cr,), generator=g, device=device) * 45000.0 + 5000.0
            f.view(-1).index_add_(0, idx, amp)
        if nan_border and k % 8 == 0:
            b = min(nan_border, rows // 4, cols // 4)
            if b:
                f[:b, :] = float("nan")
                f[-b:, :] = float("nan")
                f[:, :b] = float("nan")
                f[:, -b:] = float("nan")
        frames.append(f.contiguous())
    return frames
