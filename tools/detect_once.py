"""Profiling target for the detector: ab_align_ref_stars_dev (normalise + detect, 120 brightest) on one 8192^2 star field."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import astroburst_b200 as ab
from astroburst_b200 import pipelines, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rng = np.random.default_rng(3)
x, y, f = synth.star_catalogue(n, n, 3000, seed=3)
img = (1000.0 + synth.render_stars(n, n, x, y, f * 20) + rng.normal(0.0, 12.0, (n, n))).astype(np.float32)
dev = torch.device("cuda:0"); ctx = ab.default_context(0)
d = torch.from_numpy(img).to(dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    xy = pipelines.align_ref_stars(d.data_ptr(), n, n, ctx, s)
    print(f"align_ref_stars {n}x{n}: {1e3 * (time.perf_counter() - t0):.2f} ms, {len(xy)} stars", flush=True)
