"""A/B timing of deep-kernel build variants (AB_LIBRARY=<so>): 256 x 512 x 4096 band, parity on 8 rows."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import astroburst_b200 as ab
from oracle import oracle
dev = torch.device("cuda:0"); ctx = ab.default_context(0)
stream = torch.cuda.current_stream().cuda_stream
n_frames, rows = 256, 512
g = torch.Generator(device=dev); g.manual_seed(1)
cube = torch.randn((n_frames, rows, 4096), generator=g, device=dev) * 12.0 + 1000.0
idx = torch.randint(0, cube.numel(), (cube.numel() // 5000,), generator=g, device=dev)
cube.view(-1)[idx] += torch.rand(idx.numel(), generator=g, device=dev) * 45000.0 + 5000.0
out = torch.empty((rows, 4096), dtype=torch.float32, device=dev)
rej = torch.zeros(1, dtype=torch.int64, device=dev)
cfg = ab.StackConfig(3.0, 3.0, 5, False)
ptrs = [cube[f].data_ptr() for f in range(n_frames)]
run = lambda: ab.stack_images_device(ptrs, rows * 4096, cfg, out.data_ptr(), rej.data_ptr(), stream, ctx)
run(); run(); torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [run() for _ in range(10)]; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 10)
band = [cube[f, :8].cpu().numpy() for f in range(n_frames)]
want, _ = oracle.stack_images(band, 3.0, 3.0, 5)
print(f"{sys.argv[1] if len(sys.argv) > 1 else 'default':14s} {best:.3f} ms  {(n_frames + 1) * rows * 4096 * 4 / best / 1e6:.0f} GB/s  parity={np.array_equal(want, out[:8].cpu().numpy())}", flush=True)
