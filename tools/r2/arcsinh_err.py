import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from astroburst_b200 import imaging as im
from oracle import oracle as orc
import importlib
sky = importlib.import_module("test_imaging_gpu").sky
d = sky(500, 600, seed=6)
for factor, gamma in ((10.0, 1.0), (50.0, 1.0), (500.0, 0.5), (1.0, 2.2), (200.0, 1.7)):
    a = im.arcsinh_stretch_with_stats(d, 900.0, 30000.0, factor, gamma).astype(np.float64)
    b = orc.arcsinh_stretch_with_stats(d, 900.0, 30000.0, factor, gamma).astype(np.float64)
    nz = b != 0
    print(f"factor {factor} gamma {gamma}: max rel {np.max(np.abs(a[nz]-b[nz])/np.abs(b[nz])):.3e}  max abs {np.max(np.abs(a-b)):.3e}  differing {(a!=b).mean():.4f}")
