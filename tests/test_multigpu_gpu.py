"""2-GPU tests (NCCL): row-band sharded stack and sharded image statistics equal the single-GPU
results.  Skipped unless the box has at least two GPUs (run with `gpurun --gpus 2`)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import astroburst_b200 as ab
    from astroburst_b200 import parallel, synth

    ctx = ab.default_context(rank)
    rows, cols, n = 2304, 2048, 16                      # 4.7 Mpx > 4,000,000 -> histogram algorithm
    frames = synth.stack_frames(n, 96, cols, seed=3, nan_border=4)
    rng = np.random.default_rng(1)
    img = rng.normal(1000.0, 12.0, (rows, cols)).astype(np.float32)
    img[:50] = 0.0
    img.reshape(-1)[rng.integers(0, rows * cols, 20000)] += rng.uniform(100, 30000, 20000).astype(np.float32)
    res = {}
    # --- sharded statistics
    r0, r1 = parallel.row_band(rows, rank, world)
    band = torch.from_numpy(img[r0:r1]).to(dev).contiguous()
    cb = parallel.make_allreduce(device=dev)
    s = torch.cuda.current_stream().cuda_stream
    st = parallel.compute_image_stats_sharded(band.data_ptr(), band.numel(), cb, stream=s, ctx=ctx)
    res["stats"] = (st.min, st.max, st.median, st.mad, st.sigma, st.mean, int(st.valid_count))
    kr = parallel.compute_image_stats_sharded(band.data_ptr(), band.numel(), cb, known_range=(5.0, 40000.0),
                                              stream=s, ctx=ctx)
    res["stats_kr"] = (kr.min, kr.max, kr.median, kr.mad, kr.sigma, int(kr.valid_count))
    # the same with every reduction on the device / on the stream (NCCL through the device callback)
    cbd = parallel.make_allreduce_dev(dev)
    sd = parallel.compute_image_stats_sharded_chained(band.data_ptr(), band.numel(), cbd, stream=s, ctx=ctx)
    res["stats_dev"] = (sd.min, sd.max, sd.median, sd.mad, sd.sigma, sd.mean, int(sd.valid_count))
    kd = parallel.compute_image_stats_sharded_chained(band.data_ptr(), band.numel(), cbd, known_range=(5.0, 40000.0),
                                                      stream=s, ctx=ctx)
    res["stats_kr_dev"] = (kd.min, kd.max, kd.median, kd.mad, kd.sigma, int(kd.valid_count))
    # and over the library's peer group: reductions inside kernels over NVLink, no NCCL
    parallel.connect_peer_group(ctx)
    sp = parallel.compute_image_stats_peer(band.data_ptr(), band.numel(), stream=s, ctx=ctx)
    res["stats_peer"] = (sp.min, sp.max, sp.median, sp.mad, sp.sigma, sp.mean, int(sp.valid_count))
    kp = parallel.compute_image_stats_peer(band.data_ptr(), band.numel(), known_range=(5.0, 40000.0), stream=s, ctx=ctx)
    res["stats_kr_peer"] = (kp.min, kp.max, kp.median, kp.mad, kp.sigma, int(kp.valid_count))
    for _ in range(20):  # many epochs: the barrier flags only ever grow
        sp2 = parallel.compute_image_stats_peer(band.data_ptr(), band.numel(), stream=s, ctx=ctx)
    res["stats_peer_again"] = (sp2.min, sp2.max, sp2.median, sp2.mad, sp2.sigma, sp2.mean, int(sp2.valid_count))
    parallel.disconnect_peer_group(ctx)
    # --- sharded stack: each rank combines its row band of all frames, counters are all-reduced
    f0, f1 = parallel.row_band(96, rank, world)
    t = [torch.from_numpy(f[f0:f1].copy()).to(dev) for f in frames]
    out = torch.empty((f1 - f0, cols), dtype=torch.float32, device=dev)
    rej = torch.zeros(1, dtype=torch.int64, device=dev)
    ab.stack_images_device([x.data_ptr() for x in t], out.numel(), ab.StackConfig(3.0, 3.0, 5, False),
                           out.data_ptr(), rej.data_ptr(), s, ctx)
    parallel.allreduce_rejected(rej)
    torch.cuda.synchronize()
    res["stack_band"] = out.cpu().numpy()
    res["rej"] = int(rej)
    # --- frame-sharded input (config 5 shape): all-to-all re-shard to row bands, then the exact kernel
    for key, nf in (("fs16", 16), ("fs40", 40)):          # register path and deep (> 32 frames) path
        fr = synth.stack_frames(nf, 64, 512, seed=11, nan_border=3)
        mine = [f for f in range(nf) if parallel.frame_owner(f, nf, world) == rank]
        local = torch.from_numpy(np.ascontiguousarray(fr[mine])).to(dev)
        band, rj, (b0, b1) = parallel.stack_frame_sharded(local, nf, 64, 512,
                                                          ab.StackConfig(3.0, 3.0, 5, False), s, ctx)
        torch.cuda.synchronize()
        res[key] = (band.cpu().numpy(), int(rj), b0, b1)
    q.put((rank, res))
    dist.destroy_process_group()


def test_sharded_equals_single_gpu(orc):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import astroburst_b200 as ab
    from astroburst_b200 import synth

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rows, cols, n = 2304, 2048, 16
    rng = np.random.default_rng(1)
    img = rng.normal(1000.0, 12.0, (rows, cols)).astype(np.float32)
    img[:50] = 0.0
    img.reshape(-1)[rng.integers(0, rows * cols, 20000)] += rng.uniform(100, 30000, 20000).astype(np.float32)
    one = ab.imaging.compute_image_stats(img)
    want = (one.min, one.max, one.median, one.mad, one.sigma)
    for r in (0, 1):
        got = out[r]["stats"]
        assert got[:5] == want and got[6] == one.valid_count
        assert abs(got[5] - one.mean) <= 1e-12 * abs(one.mean)
    o = orc.compute_image_stats_with_known_range(img, 5.0, 40000.0)
    assert out[0]["stats_kr"] == out[1]["stats_kr"] == (o.min, o.max, o.median, o.mad, o.sigma, o.valid_count)
    for r in (0, 1):
        got = out[r]["stats_dev"]
        assert got[:5] == want and got[6] == one.valid_count and abs(got[5] - one.mean) <= 1e-12 * abs(one.mean)
        assert out[r]["stats_kr_dev"] == (o.min, o.max, o.median, o.mad, o.sigma, o.valid_count)
        got = out[r]["stats_peer"]
        assert got[:5] == want and got[6] == one.valid_count and abs(got[5] - one.mean) <= 1e-12 * abs(one.mean)
        assert out[r]["stats_kr_peer"] == (o.min, o.max, o.median, o.mad, o.sigma, o.valid_count)
        assert out[r]["stats_peer_again"] == out[r]["stats_peer"]
    assert out[0]["stats_peer"] == out[1]["stats_peer"]   # rank-ordered sums: identical bits on every rank
    frames = synth.stack_frames(n, 96, cols, seed=3, nan_border=4)
    want_img, want_rej = orc.stack_images(list(frames), 3.0, 3.0, 5)
    got_img = np.concatenate([out[0]["stack_band"], out[1]["stack_band"]])
    assert np.array_equal(got_img, want_img)
    assert out[0]["rej"] == out[1]["rej"] == want_rej
    for key, nf in (("fs16", 16), ("fs40", 40)):
        fr = synth.stack_frames(nf, 64, 512, seed=11, nan_border=3)
        want_img, want_rej = orc.stack_images(list(fr), 3.0, 3.0, 5)
        got = np.concatenate([out[0][key][0], out[1][key][0]])
        assert (out[0][key][2], out[0][key][3], out[1][key][2], out[1][key][3]) == (0, 32, 32, 64)
        assert np.array_equal(got, want_img)
        assert out[0][key][1] == out[1][key][1] == want_rej


def _p2p_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import astroburst_b200 as ab
    from astroburst_b200 import parallel, synth

    ctx = ab.default_context(rank)
    s = torch.cuda.current_stream().cuda_stream
    res = {}
    # register path (16), deep path (40, 256 frames), ragged band width (cols not a multiple of 32)
    for key, nf, rows, cols in (("fs16", 16, 64, 512), ("fs40", 40, 64, 512), ("fs256", 256, 10, 330)):
        fr = synth.stack_frames(nf, rows, cols, seed=11, nan_border=3)
        mine = [f for f in range(nf) if parallel.frame_owner(f, nf, world) == rank]
        local = parallel.DeviceCube((len(mine), rows, cols), ctx)
        local.tensor().copy_(torch.from_numpy(np.ascontiguousarray(fr[mine])))
        counter = parallel.DeviceCube((2,), ctx)            # 8 bytes = one u64 counter
        counter.tensor().zero_()
        cubes = parallel.PeerCubes(local)                   # (barrier inside: counters are zero everywhere)
        counters = parallel.PeerCubes(counter)
        r0, r1 = parallel.row_band(rows, rank, world)
        out = torch.empty((r1 - r0, cols), dtype=torch.float32, device=dev)
        parallel.stack_frame_sharded_p2p(cubes, nf, rows, cols, ab.StackConfig(3.0, 3.0, 5, False),
                                         out.data_ptr(), rejected_targets=counters.base, stream=s)
        band_cube = torch.empty((nf, r1 - r0, cols), dtype=torch.float32, device=dev)
        parallel.reshard_frames_to_bands_p2p(cubes, nf, rows, cols, band_cube.data_ptr(), stream=s)
        torch.cuda.synchronize()
        dist.barrier()                                      # every rank's fan-out has landed
        rej = int(counter.tensor().view(torch.int64)[0].item())
        res[key] = (out.cpu().numpy(), rej, bool(np.array_equal(band_cube.cpu().numpy(), fr[:, r0:r1], equal_nan=True)))
        counters.close()
        cubes.close()
        counter.free()
        local.free()
    q.put((rank, res))
    dist.destroy_process_group()


def test_frame_sharded_stack_with_peer_memory_loads(orc):
    """The frame-sharded stack reading its row bands straight from the owners' HBM over NVLink (one
    kernel, no staging copy; reject counts fanned out with peer atomics) equals the oracle on the
    register and deep paths; the pull re-shard behind the C ABI reproduces the bands."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from astroburst_b200 import synth

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_p2p_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        out = dict(q.get(timeout=180) for _ in procs)
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    for key, nf, rows, cols in (("fs16", 16, 64, 512), ("fs40", 40, 64, 512), ("fs256", 256, 10, 330)):
        fr = synth.stack_frames(nf, rows, cols, seed=11, nan_border=3)
        want_img, want_rej = orc.stack_images(list(fr), 3.0, 3.0, 5)
        got = np.concatenate([out[0][key][0], out[1][key][0]])
        assert np.array_equal(got, want_img), key
        assert out[0][key][1] == out[1][key][1] == want_rej, key
        assert out[0][key][2] and out[1][key][2], key
