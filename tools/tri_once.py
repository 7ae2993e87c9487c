"""Profiling target for triangle matching: match_stars on two 120-star lists of an 8192^2 field."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from astroburst_b200 import alignment
rng = np.random.default_rng(11)
ref = rng.uniform(20, 8170, (120, 2))
tgt = (ref * 1.0004 + np.array([5.5, -3.25]) + rng.normal(0, 0.08, ref.shape))[rng.permutation(120)]
for mode in ("bucketed", "all-pairs"):
    if mode == "all-pairs":
        os.environ["AB_TRI_BRUTE"] = "1"
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); m = alignment.match_stars(ref, tgt); best = min(best, time.perf_counter() - t0)
    print(f"{mode}: {best * 1e3:.3f} ms per match_stars call (host lists in, matches out), {len(m)} matches", flush=True)
