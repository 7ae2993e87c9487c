"""bench_configs.py -- BASELINE configs 3 and 4 as pipelines, for bench.py's `secondary` object.

config 3: star detect + affine align + bicubic resample, 16 frames of 8192^2, frames spread over
          the ranks (2 per GPU at 8), the reference frame's star list broadcast.
config 4: 3-channel blend + white balance + SCNR + linked STF + tone curve on 3 x 13759 x 12451,
          row-banded over the ranks, image statistics reduced with NCCL on the device.

Both time whole pipelines (wall clock between device synchronisations, max over ranks, inputs
resident in HBM) and check the results against the CPU oracle (test infrastructure) on samples.
Can also be run on its own:  python bench_configs.py [--gpus N]  (torchrun for N > 1).
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


def _sync(torch, dist, world):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _wall_max(torch, dist, world, dev, fn, reps):
    """best-of-reps wall clock of fn() between synchronisations, max over ranks (ms)."""
    best = None
    for _ in range(reps):
        _sync(torch, dist, world)
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        ms = float(dt.item()) * 1e3
        best = ms if best is None else min(best, ms)
    return best


# ------------------------------------------------------------------------------------------ config 4
C4_WEIGHTS = [(0, 0.8, 0.1, 0.0), (1, 0.2, 0.7, 0.15), (2, 0.0, 0.2, 0.85)]            # hubble_legacy-like 3x3
C4_CURVE = [(0.0, 0.0), (0.25, 0.15), (0.5, 0.5), (0.75, 0.85), (1.0, 1.0)]            # curves.rs:256-258


def c4_band(torch, dev, k, r0, r1, rows, cols):
    """Rows [r0, r1) of input channel k: smooth nebula + noise, zero padding strip at the top
    (5 % of the rows), values > 0 elsewhere.  Row-seeded: any rank can make any band."""
    yy = torch.arange(r0, r1, device=dev, dtype=torch.float32)[:, None]
    xx = torch.arange(cols, device=dev, dtype=torch.float32)[None, :]
    neb = 600.0 + 500.0 * torch.sin(yy * (0.0007 * (k + 1))) * torch.cos(xx * (0.0011 * (k + 2)))
    g = torch.Generator(device=dev)
    g.manual_seed(4000 + 7 * k + 100003 * r0)
    p = neb + torch.randn((r1 - r0, cols), generator=g, device=dev) * 30.0 + 200.0 * k
    p.clamp_(min=1.0)
    pad = rows // 20
    if r0 < pad:
        p[: min(pad, r1) - r0] = 0.0
    return p.contiguous()


def run_config4(torch, dist, ab, ctx, dev, rank, world, reps=3, rows=13759, cols=12451):
    from astroburst_b200 import parallel, pipelines
    from oracle import chains, oracle

    stream = torch.cuda.current_stream().cuda_stream
    r0, r1 = parallel.row_band(rows, rank, world)
    n = (r1 - r0) * cols
    total = rows * cols
    src = [c4_band(torch, dev, k, r0, r1, rows, cols) for k in range(3)]
    out = [torch.empty((r1 - r0, cols), dtype=torch.float32, device=dev) for _ in range(3)]
    ms_nccl = None
    res_box = [None]

    # Three lanes (host threads with their own context + stream) compute the R, G, B statistics of a
    # stage side by side (AB_C4_LANES=1: one after the other): every statistic ends in a read-back,
    # and at N > 1 in four peer reductions, which the other channels' sweeps overlap.
    from concurrent.futures import ThreadPoolExecutor

    # One GPU only: with a peer group per lane, barrier kernels of different lanes spin on the GPU at the
    # same time, and two streams that share a hardware queue then order one lane's sweep behind another
    # lane's barrier -- across ranks that is a cycle (measured: 1 in 5 runs timed out at N = 2).
    use_lanes = world == 1 and os.environ.get("AB_C4_LANES", "3") != "1"
    lane_ctx = [ctx] + ([ab.Context(dev.index), ab.Context(dev.index)] if use_lanes else [])
    # every lane on a stream of its own (never the NULL stream: its implicit ordering against other
    # streams would couple the lanes, and across ranks the peer barriers could then wait in a cycle)
    lane_stream_objs = [torch.cuda.Stream(device=dev) for _ in range(3 if use_lanes else 1)]   # (kept alive)
    lane_streams = [o.cuda_stream for o in lane_stream_objs]
    stream = lane_streams[0]
    torch.cuda.synchronize()   # the source planes (made on the current stream) exist before the lanes read them
    pool = ThreadPoolExecutor(3) if use_lanes else None

    def chain(stats_fns):
        lanes = [pipelines.StatsLane(c, st, fn) for c, st, fn in zip(lane_ctx, lane_streams, stats_fns)] if use_lanes else None
        res_box[0] = pipelines.color_chain([s.data_ptr() for s in src], n, r1 - r0, cols, C4_WEIGHTS, C4_CURVE,
                                           out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), stats_fns[0],
                                           total_pixels=total, ctx=ctx, stream=stream, lanes=lanes, pool=pool)

    if world > 1:
        # statistics reduced (a) by NCCL through the device callback (one lane: NCCL calls from several
        # host threads would have to be ordered across ranks), (b) by the library's own kernels over
        # peer memory (NVLink) -- (b) is the product path, (a) is timed beside it
        cb = parallel.make_allreduce_dev(dev)
        fn_nccl = pipelines.stats_sharded(ctx, stream, cb, total)
        lanes_saved, use_lanes = use_lanes, False
        chain([fn_nccl])
        ms_nccl = _wall_max(torch, dist, world, dev, lambda: chain([fn_nccl]), max(reps - 1, 1))
        use_lanes = lanes_saved
        for c in lane_ctx:
            parallel.connect_peer_group(c)
        stats_fns = [pipelines.stats_peer(c, st, total) for c, st in zip(lane_ctx, lane_streams)]
    else:
        stats_fns = [pipelines.stats_single_gpu(c, st) for c, st in zip(lane_ctx, lane_streams)]
    chain(stats_fns)  # warm-up
    ms = _wall_max(torch, dist, world, dev, lambda: chain(stats_fns), reps)
    res = res_box[0]
    if world > 1:
        for c in lane_ctx:
            parallel.disconnect_peer_group(c)
    if pool is not None:
        pool.shutdown()
    for c in lane_ctx[1:]:
        c.close()
    algo = 20.56e9 * (total / (13759 * 12451))  # 20.56 GB at the full size
    peak = _peak()
    info = {"workload": f"3-channel blend + WB + SCNR + linked STF + curve, {rows}x{cols} f32, row-banded over "
                        f"{world} GPU(s); 7 image statistics (3 blend, 3 WB known-range, G re-stat)"
                        + (" reduced inside kernels over peer memory (NVLink), no NCCL" if world > 1 else "")
                        + (", the three channel statistics of a stage on three host threads / contexts" if use_lanes else ""),
            "ms_total": ms, "value": total / (ms * 1e-3) / 1e6, "unit": "Mpix/s",
            "algorithmic_bytes": int(algo), "ideal_ms_at_measured_peak": algo / peak / 1e6 / world,
            "frac_of_roofline": algo / (ms * 1e-3) / 1e9 / (peak * world), "gpu_launches": int(res.launches),
            "timing": "wall clock between device synchronisations, max over ranks, best of reps"}
    if ms_nccl is not None:
        info["ms_total_with_nccl_callback_stats"] = ms_nccl
    # ---- parity: (a) statistics of the blended R plane, whole image, against the oracle; (b) final
    # pixels of sampled rows of every rank's band with the oracle chain on the same statistics
    okt = torch.ones(1, dtype=torch.int32, device=dev)
    chk = 4
    s_b, s_w = res.stats_blend, res.stats_balanced

    def ost(st):
        return oracle.ImageStats(st.min, st.max, st.median, st.mad, st.sigma, st.mean, st.valid_count)

    mid = (r1 - r0) // 2  # (the first rows of the image are the zero padding strip)
    rows_in = [s[mid:mid + chk].cpu().numpy() for s in src]
    oracle.set_num_threads(max((os.cpu_count() or 1) // world, 1))
    want, _, _, ostf = chains.color_chain(rows_in, C4_WEIGHTS, C4_CURVE, stats_blend=tuple(ost(x) for x in s_b),
                                          stats_balanced=tuple(ost(x) for x in s_w))
    got = [o[mid:mid + chk].cpu().numpy() for o in out]
    exact = all(np.array_equal(a, b) for a, b in zip(got, want))
    rel = max(float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6))) for a, b in zip(got, want))
    if not (exact or rel <= 1e-5):
        okt.zero_()
    if (ostf.shadow, ostf.midtone, ostf.highlight) != (res.stf.shadow, res.stf.midtone, res.stf.highlight):
        okt.zero_()
    # (a): the blended R plane gathered on rank 0 -> oracle.compute_image_stats of the whole image
    blend_r = torch.empty((r1 - r0, cols), dtype=torch.float32, device=dev)
    w = (ab._lib.BlendWeightC * 3)(*[ab._lib.BlendWeightC(*x) for x in C4_WEIGHTS])
    ptrs = (C.c_void_p * 3)(*[s.data_ptr() for s in src])
    lens = (C.c_size_t * 3)(n, n, n)
    tmp_g, tmp_b = torch.empty_like(blend_r), torch.empty_like(blend_r)
    ctx.check(ab.lib().ab_blend_channels_dev(ctx.handle, ptrs, lens, 3, w, 3, r1 - r0, cols, blend_r.data_ptr(),
                                             tmp_g.data_ptr(), tmp_b.data_ptr(), stream))
    torch.cuda.synchronize()
    del tmp_g, tmp_b
    host_full = None
    if world > 1:
        if rank == 0:
            host_full = np.empty((rows, cols), np.float32)
            host_full[r0:r1] = blend_r.cpu().numpy()
            for q in range(1, world):
                q0, q1 = parallel.row_band(rows, q, world)
                buf = torch.empty((q1 - q0, cols), dtype=torch.float32, device=dev)
                dist.recv(buf, src=q)
                host_full[q0:q1] = buf.cpu().numpy()
                del buf
        else:
            dist.send(blend_r, dst=0)
    else:
        host_full = blend_r.cpu().numpy()
    stats_ok = None
    if rank == 0:
        oracle.set_num_threads(os.cpu_count() or 1)
        o = oracle.compute_image_stats(host_full)
        g = s_b[0]
        stats_ok = (g.min, g.max, g.median, g.mad, g.sigma, int(g.valid_count)) == (
            o.min, o.max, o.median, o.mad, o.sigma, int(o.valid_count)) and abs(g.mean - o.mean) <= 1e-12 * abs(o.mean)
        if not stats_ok:
            okt.zero_()
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    info["matches_oracle"] = bool(okt.item())
    info["oracle_check"] = (f"{chk} rows of every rank's band through the whole chain "
                            f"({'bit-exact' if exact else f'max rel err {rel:.2e}'} on rank 0) + ImageStats of the "
                            f"blended R plane ({rows}x{cols}) against the CPU oracle")
    info["wb_factors"] = list(res.factors)
    del src, out, blend_r
    torch.cuda.empty_cache()
    return info


# ------------------------------------------------------------------------------------------ config 3
def c3_catalogue(R):
    rng = np.random.default_rng(3)
    nst = int(3000 * (R / 4096) ** 2)
    x, y = rng.uniform(20, R - 20, nst), rng.uniform(20, R - 20, nst)
    fl = 300.0 * (40000.0 / 300.0) ** (rng.uniform(0, 1, nst) ** 2)
    return x, y, fl


def c3_transform(f):
    """known affine of frame f (frame 0 = identity): rotation <= 2 deg, |t| <= 40 px, scale 1 +- 0.01"""
    if f == 0:
        return (1.0, 0.0, 0.0, 0.0, 1.0, 0.0)
    rng = np.random.default_rng(300 + f)
    th = math.radians(rng.uniform(-2.0, 2.0))
    sc = 1.0 + rng.uniform(-0.01, 0.01)
    return (sc * math.cos(th), -sc * math.sin(th), rng.uniform(-40, 40), sc * math.sin(th), sc * math.cos(th),
            rng.uniform(-40, 40))


def c3_frame(torch, dev, f, R, cat):
    x, y, fl = cat
    t = c3_transform(f)
    g3 = torch.Generator(device=dev)
    g3.manual_seed(3000 + f)
    img = torch.randn((R, R), generator=g3, device=dev) * 4.0 + 200.0
    xs = torch.tensor(t[0] * x + t[1] * y + t[2], device=dev)
    ys = torch.tensor(t[3] * x + t[4] * y + t[5], device=dev)
    fv = torch.tensor(fl, device=dev)
    rr = 7
    dy, dx = torch.meshgrid(torch.arange(-rr, rr + 1, device=dev), torch.arange(-rr, rr + 1, device=dev), indexing="ij")
    cx, cy = xs.round().long(), ys.round().long()
    px, py = cx[:, None, None] + dx[None], cy[:, None, None] + dy[None]
    val = fv[:, None, None] / (2 * math.pi * 1.5 ** 2) * torch.exp(
        -((px - xs[:, None, None]) ** 2 + (py - ys[:, None, None]) ** 2) / (2 * 1.5 ** 2))
    ok = (px >= 0) & (px < R) & (py >= 0) & (py < R)
    img.view(-1).index_add_(0, (py * R + px)[ok], val[ok].float())
    return img.contiguous()


def run_config3(torch, dist, ab, ctx, dev, rank, world, reps=2, n_frames=16, R=8192):
    from astroburst_b200 import parallel, pipelines
    from oracle import oracle

    stream = torch.cuda.current_stream().cuda_stream
    cat = c3_catalogue(R)
    mine = [f for f in range(n_frames) if parallel.frame_owner(f, n_frames, world) == rank]
    frames = {f: c3_frame(torch, dev, f, R, cat) for f in mine}
    outs = {f: torch.empty((R, R), dtype=torch.float32, device=dev) for f in mine if f != 0}
    results = {}
    xy_dev = torch.zeros((121, 2), dtype=torch.float64, device=dev)
    # Three host threads feed the GPU (AB_C3_LANES; measured 54.2 / 46.4 / 40.5 / 40.9 ms for 1 / 2 / 3 / 4 on
    # one B200 with 16 frames), each with its own context and streams -- pipelines.align_channels with the
    # reference list coming from another rank:
    # the detector synchronises a few times per frame (counts that size the next launch), and the
    # other lane's kernels fill those gaps.  Entry points release the GIL (ctypes) and are
    # thread-safe per context.
    from concurrent.futures import ThreadPoolExecutor

    n_lanes = max(1, min(int(os.environ.get("AB_C3_LANES", "3")), 4))
    ctxs = [ctx] + [ab.Context(dev.index) for _ in range(n_lanes - 1)]
    lanes = [torch.cuda.Stream(device=dev) for _ in range(n_lanes)]
    warp_sides = [torch.cuda.Stream(device=dev) for _ in range(n_lanes)]   # warps overlap the next frame's matching
    targets = [f for f in mine if f != 0]
    pool = ThreadPoolExecutor(n_lanes)

    def pipeline():
        # reference star list once, on the rank that holds frame 0, broadcast (<= 1.9 KB) to the
        # others -- which detect the stars of their own target frames while it is on its way
        def detect_lane(k):
            return {f: pipelines.align_ref_stars(frames[f].data_ptr(), R, R, ctxs[k], lanes[k].cuda_stream)
                    for f in targets[k::n_lanes]}

        futs = [pool.submit(detect_lane, k) for k in range(1, n_lanes)]
        work = None
        if rank == 0:
            xy = pipelines.align_ref_stars(frames[0].data_ptr(), R, R, ctx, lanes[0].cuda_stream)
            xy_dev.zero_()
            xy_dev[0, 0] = len(xy)
            if len(xy):
                xy_dev[1:1 + len(xy)].copy_(torch.from_numpy(xy))
        if world > 1:
            work = dist.broadcast(xy_dev, src=0, async_op=True)
        tgt_xy = detect_lane(0)
        for fu in futs:
            tgt_xy.update(fu.result())
        if work is not None:
            work.wait()
        h = xy_dev.cpu().numpy()
        ref_xy = h[1:1 + int(h[0, 0])]
        ref_ptr = frames[0].data_ptr() if 0 in frames else None

        def finish_lane(k):
            out = {}
            for f in targets[k::n_lanes]:
                out[f] = pipelines.align_and_warp_from_stars(ref_xy, tgt_xy[f], ref_ptr, frames[f].data_ptr(),
                                                             outs[f].data_ptr(), R, R, 16, ctxs[k],
                                                             lanes[k].cuda_stream,
                                                             warp_stream=warp_sides[k].cuda_stream)
            return out

        futs = [pool.submit(finish_lane, k) for k in range(1, n_lanes)]
        results.update(finish_lane(0))
        for fu in futs:
            results.update(fu.result())

    torch.cuda.synchronize()   # the frames exist before the lanes' streams touch them
    pipeline()
    ms = _wall_max(torch, dist, world, dev, pipeline, reps)
    pool.shutdown()
    for c in ctxs[1:]:
        c.close()
    npix = R * R
    algo = (2 * npix * 4) + (n_frames - 1) * (4 * npix * 4)  # ref 2 reads + 15 x (2 reads + warp R+W)
    peak = _peak()
    # ---- checks: recovered transforms against the known ones (corner error), warp rows against the oracle
    okt = torch.ones(1, dtype=torch.int32, device=dev)
    worst = 0.0
    methods = []
    for f, res in results.items():
        got = [getattr(res.transform, k) for k in "a b tx c d ty".split()]
        t = c3_transform(f)
        err = max(math.hypot(got[0] * X + got[1] * Y + got[2] - (t[0] * X + t[1] * Y + t[2]),
                             got[3] * X + got[4] * Y + got[5] - (t[3] * X + t[4] * Y + t[5]))
                  for X, Y in ((0, 0), (R, 0), (0, R), (R, R)))
        worst = max(worst, err)
        methods.append(int(res.method))
        if not (err < 1.0 and res.method in (0, 1)):
            okt.zero_()
    warp_ok = None
    if results:
        f = sorted(results)[0]
        got = tuple(getattr(results[f].transform, k) for k in "a b tx c d ty".split())
        oracle.set_num_threads(max((os.cpu_count() or 1) // world, 1))
        want = oracle.warp_image(frames[f].cpu().numpy(), got, 64, R)
        warp_ok = bool(np.array_equal(want, outs[f][:64].cpu().numpy()))
        if not warp_ok:
            okt.zero_()
    wt = torch.tensor([worst], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        dist.all_reduce(wt, op=dist.ReduceOp.MAX)
    info = {"workload": f"star detect + affine align + bicubic warp, {n_frames} frames {R}x{R} f32, "
                        f"{n_frames // world if n_frames >= world else 1} frame(s) per GPU over {world} GPU(s), "
                        f"reference star list broadcast, {n_lanes} host thread(s) per GPU",
            "ms_total": ms, "value": (n_frames - 1) * npix / (ms * 1e-3) / 1e6, "unit": "Mpix/s (aligned + warped)",
            "ms_per_target_frame_per_gpu": ms / max(len([f for f in mine if f != 0]), 1),
            "algorithmic_bytes": int(algo), "ideal_ms_at_measured_peak": algo / peak / 1e6 / world,
            "frac_of_roofline": algo / (ms * 1e-3) / 1e9 / (peak * world),
            "matches_oracle": bool(okt.item()), "max_corner_error_px": float(wt.item()),
            "oracle_check": "every recovered transform within 1 px of the known one at the corners (methods "
                            f"{sorted(set(methods))} on rank 0); 64 warped rows bit-exact against the CPU oracle per rank",
            "timing": "wall clock between device synchronisations, max over ranks, best of reps"}
    del frames, outs
    torch.cuda.empty_cache()
    return info


def main():
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--small", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import astroburst_b200 as ab

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = ab.default_context(local)
    kw4 = dict(rows=3440, cols=3113) if args.small else {}
    kw3 = dict(R=2048, n_frames=max(4, world * 2)) if args.small else {}
    out = {"config4": run_config4(torch, dist, ab, ctx, dev, rank, world, **kw4),
           "config3": run_config3(torch, dist, ab, ctx, dev, rank, world, **kw3)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
