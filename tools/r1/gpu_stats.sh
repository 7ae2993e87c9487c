#!/bin/bash
# Statistics kernels: parity tests, per-plane timing, ncu capture of the four sweeps; deep-stack timing.
mkdir -p gpurun_out
python -m pytest tests/test_imaging_gpu.py tests/test_stack_gpu.py tests/test_calibration_gpu.py tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -8
python tools/stats_once.py 2>&1 | tee gpurun_out/stats_once.txt
ncu --set full --clock-control none --import-source on -k regex:"hist_kernel|scan_kernel" -c 4 -f -o gpurun_out/prof_stats python tools/stats_once.py > gpurun_out/prof_stats.log 2>&1
python - <<'PY' 2>&1 | tee gpurun_out/deep.txt
import numpy as np, torch
import astroburst_b200 as ab
from astroburst_b200 import synth
from oracle import oracle
dev = torch.device("cuda:0"); ctx = ab.default_context(0)
stream = torch.cuda.current_stream().cuda_stream
for n_frames, rows in ((256, 512), (128, 512), (64, 512)):
    fr = synth.stack_frames_torch(n_frames, rows, 4096, dev, seed=5)
    cube = torch.stack(fr); del fr
    out = torch.empty((rows, 4096), dtype=torch.float32, device=dev)
    rej = torch.zeros(1, dtype=torch.int64, device=dev)
    cfg = ab.StackConfig(3.0, 3.0, 5, False)
    ptrs = [cube[f].data_ptr() for f in range(n_frames)]
    run = lambda: ab.stack_images_device(ptrs, rows * 4096, cfg, out.data_ptr(), rej.data_ptr(), stream, ctx)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [run() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    band = [cube[f, :8].cpu().numpy() for f in range(n_frames)]
    want, wrej = oracle.stack_images(band, 3.0, 3.0, 5)
    got = out[:8].cpu().numpy()
    nbytes = (n_frames + 1) * rows * 4096 * 4
    print(f"n={n_frames} rows={rows}: {ms:.3f} ms  {nbytes/ms/1e6:.0f} GB/s  parity(8 rows)={np.array_equal(want, got)} rej={int(rej)}")
    del cube
PY
