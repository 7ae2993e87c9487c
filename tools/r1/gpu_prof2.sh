#!/bin/bash
# Star-detector parity after the host flood-fill rewrite, per-operator bench, ncu of the deep-stack and gather kernels.
mkdir -p gpurun_out
python -m pytest tests/test_stars_gpu.py tests/test_preview_gpu.py tests/test_phasecorr_gpu.py -m gpu -x -q 2>&1 | tail -5
python bench_ops.py --out gpurun_out/bench_ops.json 2>&1 | grep -E "detect_stars|align_channel|warp_image|masked_stretch" | cut -c1-400
ncu --set full --clock-control none --import-source on -k regex:stack_deep -c 1 -f -o gpurun_out/prof_deep2 python tools/deep_once.py 256 128 > gpurun_out/prof_deep2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gather_kernel -c 1 -f -o gpurun_out/prof_gather python - > gpurun_out/prof_gather.log 2>&1 <<'PY'
import numpy as np, torch, ctypes as C, math
import astroburst_b200 as ab
from astroburst_b200 import _lib
from astroburst_b200._lib import lib
dev = torch.device("cuda:0"); ctx = ab.default_context(0)
R = 8192
img = torch.randn((R, R), device=dev) * 4 + 200
out = torch.empty_like(img)
th = math.radians(1.1)
t = _lib.AffineC(1.004 * math.cos(th), -math.sin(th), 23.5, math.sin(th), 1.004 * math.cos(th), -17.25)
ctx.check(lib().ab_warp_affine_bicubic_dev(ctx.handle, img.data_ptr(), R, R, C.addressof(t), R, R, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
PY
ls -la gpurun_out | tail -5
