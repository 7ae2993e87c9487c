"""Profiling target for the tile background kernel: estimate_background on one 8192^2 star field.
python tools/tiles_once.py [size]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from astroburst_b200 import alignment, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rng = np.random.default_rng(3)
x, y, f = synth.star_catalogue(n, n, 2000, seed=3)
img = (1000.0 + synth.render_stars(n, n, x, y, f) + rng.normal(0.0, 12.0, (n, n))).astype(np.float32)
for _ in range(3):
    t0 = time.perf_counter()
    med, sig = alignment.estimate_background(img, 256)
    print(f"estimate_background {n}x{n}: {1e3 * (time.perf_counter() - t0):.2f} ms (host buffers), median {med:.4f} sigma {sig:.4f}", flush=True)
