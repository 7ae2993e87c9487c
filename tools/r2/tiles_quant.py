"""Tile background kernel on tie-heavy planes (integer ADU values): fast-path hit rate and time."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from astroburst_b200 import alignment, synth
from oracle import oracle
n = 4096
rng = np.random.default_rng(3)
x, y, f = synth.star_catalogue(n, n, 800, seed=3)
base = 1000.0 + synth.render_stars(n, n, x, y, f) + rng.normal(0.0, 12.0, (n, n))
for name, img in (("continuous", base.astype(np.float32)), ("integer ADU", np.rint(base).astype(np.float32)),
                  ("integer ADU / 65535", (np.rint(base) / 65535.0).astype(np.float32)),
                  ("8 levels", (np.rint(base / 8.0) * 8.0).astype(np.float32))):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); med, sig = alignment.estimate_background(img, 256); best = min(best, time.perf_counter() - t0)
    om, osg = oracle.estimate_background(img, 256)
    print(f"{name:22s}: {best * 1e3:7.2f} ms (host plane in), median {med!r} sigma {sig!r} oracle-equal={med == om and sig == osg}", flush=True)
