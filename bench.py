#!/usr/bin/env python
"""bench.py -- headline benchmark: Mpix/s of the sigma-clip stack (32 x 4096^2 f32).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (stack_images core loop, combine.rs:160-182) over one
32-frame 4096x4096 f32 synthetic stack per GPU.  `value` = output pixels of all ranks / the
max-over-ranks device time (inputs resident in HBM).  `e2e` = the same through the host-buffer
C-ABI call (pinned host frames -> H2D -> kernel -> D2H) inside the timed region.
`--impl reference` times the CPU oracle (restatement of the reference's Rust; no Rust
toolchain here) on a bounded row-band sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FRAMES, ROWS, COLS = 32, 4096, 4096
SIGMA_LOW, SIGMA_HIGH, MAX_ITER = 3.0, 3.0, 5
ALGO_BYTES = N_FRAMES * ROWS * COLS * 4 + ROWS * COLS * 4  # 2,214,592,512 B
METRIC = "Mpix/s sigma-clip stack (32x4096^2 f32)"
WORKLOAD = "32-frame sigma-clip stack, 4096x4096 f32, sigma 3/3, 5 iters, align=false"


def _measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one stack launch, from the committed
    `ncu --set full` capture of this command (profiles/stack_ncu_traffic.json), else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "stack_ncu_traffic.json")) as f:
            return int(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU during the timed region (NVML in a thread,
    1 ms period; falls back to `nvidia-smi -lms` when the NVML binding is missing)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.sm = []
        self.reasons = set()
        self.mx = None
        self.stop_flag = False
        self.nvml = None

    def _nvml_loop(self):
        n, h = self.nvml, self.handle
        bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
                for nm, b in bits.items():
                    if r & b:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        try:
            import pynvml as n
            n.nvmlInit()
            # NVML indices follow PCI order; CUDA_VISIBLE_DEVICES may remap -> go through the UUID
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            try:
                self.handle = n.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.handle = n.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
            self.nvml = n
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                    "sm_max_mhz": self.mx, "samples": len(self.sm),
                    "reasons": sorted(self.reasons), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            parts = [x.strip() for x in l.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def cpu_sample_rate(frames_rows, threads=None):
    """Oracle (CPU port of the reference) on a row-band sample.  -> (Mpix/s, seconds, cores)."""
    from oracle import oracle

    if threads:
        oracle.set_num_threads(threads)
    cores = oracle.num_threads()
    oracle.stack_images([f[:8] for f in frames_rows], SIGMA_LOW, SIGMA_HIGH, MAX_ITER)  # warm
    t0 = time.perf_counter()
    out, rej = oracle.stack_images(frames_rows, SIGMA_LOW, SIGMA_HIGH, MAX_ITER)
    dt = time.perf_counter() - t0
    return out.size / dt / 1e6, dt, cores, out, rej


def run_reference(args):
    """CPU arm: the oracle port on all host threads, each step a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from astroburst_b200 import synth

    sample_rows = 512
    frames = list(synth.stack_frames(N_FRAMES, sample_rows, COLS, seed=42, nan_border=0))
    for _ in range(max(args.warmup, 1)):
        cpu_sample_rate([f[:16] for f in frames])
    rates, t_tot = [], 0.0
    cores = 0
    for _ in range(args.steps):
        r, dt, cores, _, _ = cpu_sample_rate(frames)
        rates.append(r)
        t_tot += dt
    value = sample_rows * COLS * args.steps / t_tot / 1e6
    sample = f"{sample_rows} of {ROWS} rows x {COLS} cols x {N_FRAMES} frames per step (same generator)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_tot / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 compares, f64 sums", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "note": "CPU oracle = C++/OpenMP restatement of the reference's Rust+rayon stack_images; "
                "the Rust reference cannot be built in this image (no rustc/cargo).",
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    import astroburst_b200 as ab
    from astroburst_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    ctx = ab.default_context(local)
    cfg = ab.StackConfig(SIGMA_LOW, SIGMA_HIGH, MAX_ITER, False)
    # per-rank independent stack (weak scaling: a mosaic N x taller, row-banded over ranks)
    # HBM layout: one [N, rows, cols] cube (constant frame stride -> one 2-D TMA per tile)
    frames = synth.stack_frames_torch(N_FRAMES, ROWS, COLS, dev, seed=42 + rank)
    cube = torch.stack(frames)
    del frames
    frames = [cube[f] for f in range(N_FRAMES)]
    ptrs = [f.data_ptr() for f in frames]
    out = torch.empty((ROWS, COLS), dtype=torch.float32, device=dev)
    rej = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    npix = ROWS * COLS

    # The path's only exchange is StackResult.rejected_pixels: one 8-byte all-reduce
    # per stack.  It is issued asynchronously on a ring of counters, so the reduction of stack k
    # overlaps the kernel of stack k+1 (a pipeline of stacks); every reduction is complete before
    # the closing synchronisation of the timed region.
    rej_ring = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(4)]
    pending = [None] * 4
    step_no = [0]

    def step():
        k = step_no[0] % 4
        step_no[0] += 1
        if pending[k] is not None:
            pending[k].wait()
            pending[k] = None
        r = rej_ring[k] if world > 1 else rej
        ab.stack_images_device(ptrs, npix, cfg, out.data_ptr(), r.data_ptr(), stream, ctx)
        if world > 1:
            pending[k] = dist.all_reduce(r, async_op=True)

    def drain():
        for k in range(4):
            if pending[k] is not None:
                pending[k].wait()
                pending[k] = None
        if world > 1:
            rej.copy_(rej_ring[(step_no[0] - 1) % 4])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    drain()
    sync_all()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    # cudaProfilerStart/Stop bracket the timed region: `ncu --profile-from-start off` then lists
    # exactly the launches of the timed steps (no synthetic-data kernels); a no-op otherwise
    torch.cuda.cudart().cudaProfilerStart()
    t_start.record()
    for k in range(args.steps):
        ev[k][0].record()
        step()
        ev[k][1].record()
    drain()
    t_end.record()
    sync_all()
    torch.cuda.cudart().cudaProfilerStop()
    launches = ctx.launch_count - launches0
    total_ms = t_start.elapsed_time(t_end)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = world * npix / (ms_per_step * 1e-3) / 1e6
    rejected = int(rej.item())

    # ---- e2e: host-buffer C-ABI call, pinned frames, H2D + kernel + D2H inside the timed region
    host = [torch.empty((ROWS, COLS), dtype=torch.float32, pin_memory=True) for _ in range(N_FRAMES)]
    for h, f in zip(host, frames):
        h.copy_(f)
    host_out = torch.empty((ROWS, COLS), dtype=torch.float32, pin_memory=True)
    host_np = [h.numpy() for h in host]
    torch.cuda.synchronize()
    res = ab.stack_images(host_np, cfg, ctx, out=host_out.numpy())  # warm (allocates scratch)
    e2e_same = bool(np.array_equal(res.image, out.cpu().numpy())) and res.rejected_pixels == (
        rejected if world == 1 else res.rejected_pixels)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        ab.stack_images(host_np, cfg, ctx, out=host_out.numpy())
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * npix * args.e2e_steps / float(te.item()) / 1e6

    if rank == 0:
        peak, peak_src = _measured_peak()
        kern_ms = float(np.mean(step_ms))
        achieved = ALGO_BYTES / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 compares, f64 sums", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu": "one 32x4096^2 stack per rank",
                       "l2": "inputs (2.1 GB/step) larger than L2, no flush needed",
                       "rejected_pixels": rejected, "input_msamples_per_s": value * N_FRAMES},
            "clocks": clocks,
            "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "Mpix/s",
                    "h2d_bytes_per_step": N_FRAMES * npix * 4, "d2h_bytes_per_step": npix * 4 + 8,
                    "steps": args.e2e_steps, "matches_device_path": e2e_same},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": _ncu_traffic(), "peak_source": peak_src,
                         "kernel": "abk::stack_wtma_kernel<32, 1, 2, true> (one launch per step)",
                         "kernel_ms": kern_ms, "algorithmic_bytes": ALGO_BYTES},
        }
        if not args.no_cpu_baseline and world == 1:
            sample_rows = ROWS  # the whole stack: ~1 s on 16 host cores (~15 core-seconds)
            r0 = 0
            band = [f[r0:r0 + sample_rows].cpu().numpy() for f in frames]
            rate, dt, cores, cpu_out, _ = cpu_sample_rate(band)
            same = bool(np.array_equal(cpu_out, out[r0:r0 + sample_rows].cpu().numpy()))
            line["cpu_baseline"] = {
                "value": rate, "unit": "Mpix/s", "cores": cores, "kind": "port",
                "sample": f"rows {r0}..{r0 + sample_rows} of the same frames ({dt:.1f} s)",
                "gpu_matches_oracle_on_sample": same}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
