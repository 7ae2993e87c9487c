"""Per-call wall times of the config-4 colour chain on N GPUs (torchrun): where the time goes."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import astroburst_b200 as ab
import bench_configs as bc
from astroburst_b200 import parallel, pipelines
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = ab.default_context(local)
rows, cols = 13759, 12451
stream = torch.cuda.current_stream().cuda_stream
r0, r1 = parallel.row_band(rows, rank, world)
n = (r1 - r0) * cols
src = [bc.c4_band(torch, dev, k, r0, r1, rows, cols) for k in range(3)]
out = [torch.empty((r1 - r0, cols), dtype=torch.float32, device=dev) for _ in range(3)]
if world > 1:
    parallel.connect_peer_group(ctx)
    base = pipelines.stats_peer(ctx, stream, rows * cols)
else:
    base = pipelines.stats_single_gpu(ctx, stream)
log = []
def stats_fn(ptr, nn, known_range=None):
    t0 = time.perf_counter(); st = base(ptr, nn, known_range); log.append(("stats" + ("_known" if known_range else ""), (time.perf_counter() - t0) * 1e3)); return st
def chain():
    return pipelines.color_chain([s.data_ptr() for s in src], n, r1 - r0, cols, bc.C4_WEIGHTS, bc.C4_CURVE, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), stats_fn, total_pixels=rows * cols, ctx=ctx, stream=stream)
for rep in range(4):
    log.clear(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter(); res = chain(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    if rank == 0 and rep == 3:
        print(f"N={world}: chain returned after {(t1 - t0) * 1e3:.3f} ms, drained after {(t2 - t0) * 1e3:.3f} ms; launches {res.launches}")
        print("  " + ", ".join(f"{k} {v:.3f}" for k, v in log), f"| sum of statistics {sum(v for _, v in log):.3f} ms")
if world > 1:
    parallel.disconnect_peer_group(ctx); dist.destroy_process_group()
