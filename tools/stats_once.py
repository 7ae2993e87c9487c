"""Per-pass timing of compute_image_stats on one config-4 plane (profiling target):
python tools/stats_once.py [rows cols]"""
import ctypes as C, sys, time, torch
sys.path.insert(0, ".")
import astroburst_b200 as ab
from astroburst_b200 import _lib
from astroburst_b200._lib import lib
rows, cols = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (13759, 12451)
dev = torch.device("cuda:0"); ctx = ab.default_context(0); h = ctx.handle
g = torch.Generator(device=dev); g.manual_seed(4)
yy = torch.arange(rows, device=dev, dtype=torch.float32)[:, None]
xx = torch.arange(cols, device=dev, dtype=torch.float32)[None, :]
p = 600.0 + 500.0 * torch.sin(yy * 0.0007) * torch.cos(xx * 0.0022) + torch.randn((rows, cols), generator=g, device=dev) * 30.0
p.clamp_(min=1.0); p[: rows // 40] = 0.0
sky = torch.full((rows, cols), 1000.0, device=dev) + torch.randn((rows, cols), generator=g, device=dev) * 12.0
sky.view(-1)[:: 9973] += 40000.0  # sky-peaked frame with bright outliers (same-bin contention case)
stream = torch.cuda.current_stream().cuda_stream
st = _lib.ImageStatsC()
for name, img in (("nebula", p), ("sky-peaked", sky)):
    for mode, label in ((0, "4 sweeps"), (3, "known range, 3 sweeps")):
        kmin, kmax = (float(img[img > 1e-7].min()), float(img.max())) if mode == 3 else (0.0, 0.0)
        best = 1e9
        for _ in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ctx.check(lib().ab_image_stats_dev(h, img.data_ptr(), img.numel(), mode, kmin, kmax, C.addressof(st), stream))
            torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t0) * 1e3)
        nb = (4 if mode == 0 else 3) * img.numel() * 4
        print(f"{name:10s} {label:22s}: {best:.3f} ms  {nb / best / 1e6:.0f} GB/s  median={st.median:.6f} mad={st.mad:.6f} n={st.valid_count}", flush=True)
