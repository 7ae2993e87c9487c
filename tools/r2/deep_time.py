"""Timing + parity of the deep-stack kernel on one GPU: python tools/r2/deep_time.py [out.txt]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import astroburst_b200 as ab
from astroburst_b200 import synth
from oracle import oracle
dev = torch.device("cuda:0"); ctx = ab.default_context(0)
stream = torch.cuda.current_stream().cuda_stream
lines = []
for n_frames, rows in ((256, 512), (225, 512), (200, 512), (128, 512), (100, 512), (65, 512), (64, 512), (63, 512), (50, 512), (48, 512), (33, 512), (32, 512)):
    fr = synth.stack_frames_torch(n_frames, rows, 4096, dev, seed=5)
    cube = torch.stack(fr); del fr
    out = torch.empty((rows, 4096), dtype=torch.float32, device=dev)
    rej = torch.zeros(1, dtype=torch.int64, device=dev)
    cfg = ab.StackConfig(3.0, 3.0, 5, False)
    ptrs = [cube[f].data_ptr() for f in range(n_frames)]
    run = lambda: ab.stack_images_device(ptrs, rows * 4096, cfg, out.data_ptr(), rej.data_ptr(), stream, ctx)
    run(); run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [run() for _ in range(5)]; e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    band = [cube[f, :8].cpu().numpy() for f in range(n_frames)]
    want, wrej = oracle.stack_images(band, 3.0, 3.0, 5)
    got = out[:8].cpu().numpy()
    nbytes = (n_frames + 1) * rows * 4096 * 4
    lines.append(f"n={n_frames} rows={rows}: {ms:.3f} ms  {nbytes/ms/1e6:.0f} GB/s  parity(8 rows)={np.array_equal(want, got)} rej={int(rej)}")
    print(lines[-1], flush=True)
    del cube
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write("\n".join(lines) + "\n")
