"""Profiling target for the bicubic warp: one 8192^2 plane through ab_warp_affine_bicubic_dev."""
import ctypes as C, sys, time, torch
sys.path.insert(0, ".")
import astroburst_b200 as ab
from astroburst_b200._lib import lib, AffineC
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0"); ctx = ab.default_context(0)
g = torch.Generator(device=dev); g.manual_seed(1)
src = torch.rand((n, n), generator=g, device=dev); out = torch.empty_like(src)
t = AffineC(0.9996, -0.0123, 14.25, 0.0121, 1.0003, -9.5)
stream = torch.cuda.current_stream().cuda_stream
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.check(lib().ab_warp_affine_bicubic_dev(ctx.handle, src.data_ptr(), n, n, C.addressof(t), n, n, out.data_ptr(), stream))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"warp {n}x{n}: {ms:.3f} ms, {2 * n * n * 4 / ms / 1e6:.0f} GB/s (read + write once)", flush=True)
