"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: band partition, the frame->band
plan and the all-reduce callback the sharded statistics call back into."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_row_band_partition():
    from astroburst_b200.parallel import row_band

    for rows in (0, 1, 7, 4096, 12451):
        for world in (1, 2, 3, 8):
            bands = [row_band(rows, r, world) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == rows
            assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
            sizes = [b[1] - b[0] for b in bands]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        row_band(10, 3, 2)


def test_frame_to_band_plan_covers_everything():
    from astroburst_b200.parallel import frame_to_band_plan, row_band

    n_frames, rows, world = 256, 4096, 8
    plan = frame_to_band_plan(n_frames, rows, world)
    seen = np.zeros((n_frames, rows), np.int32)
    per_src = [0] * world
    for src in range(world):
        for dst in range(world):
            for f, r0, r1 in plan[src][dst]:
                assert (r0, r1) == row_band(rows, dst, world)
                seen[f, r0:r1] += 1
                per_src[src] += 1
    assert (seen == 1).all()
    assert len(set(per_src)) == 1       # 32 frames per rank (config 5)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from astroburst_b200.parallel import allreduce_rejected, make_allreduce

    cb = make_allreduce()
    res = {}
    # what libastroburst_cuda does: hand the callback a raw host buffer
    hist = np.arange(65536, dtype=np.uint64) * np.uint64(rank + 1)
    assert cb(None, hist.ctypes.data, hist.size, 1) == 0
    res["hist_ok"] = bool((hist == np.arange(65536, dtype=np.uint64) * np.uint64(3)).all())
    s = np.array([1.5 + rank], np.float64)
    cb(None, s.ctypes.data, 1, 0)
    res["sum"] = float(s[0])
    mn = np.array([10.0 - rank], np.float64)
    mx = np.array([10.0 + rank], np.float64)
    cb(None, mn.ctypes.data, 1, 2)
    cb(None, mx.ctypes.data, 1, 3)
    res["min"], res["max"] = float(mn[0]), float(mx[0])
    rej = torch.tensor([100 + rank], dtype=torch.int64)
    res["rej"] = int(allreduce_rejected(rej))
    # sharded statistics by the reference algorithm on CPU tensors, to pin the data flow the GPU
    # path follows: band-local histograms summed over ranks == histogram of the whole image
    rng = np.random.default_rng(0)
    img = rng.normal(1000, 12, (64, 128)).astype(np.float32)
    from astroburst_b200.parallel import row_band
    r0, r1 = row_band(64, rank, world)
    band = img[r0:r1]
    lo, hi = float(img.min()), float(img.max())
    idx = np.minimum(((band.astype(np.float64) - lo) * (65536 / (hi - lo))).astype(np.int64), 65535)
    h = np.bincount(idx.reshape(-1), minlength=65536).astype(np.uint64)
    cb(None, h.ctypes.data, h.size, 1)
    full = np.minimum(((img.astype(np.float64) - lo) * (65536 / (hi - lo))).astype(np.int64), 65535)
    res["band_hist_ok"] = bool((h == np.bincount(full.reshape(-1), minlength=65536).astype(np.uint64)).all())
    q.put((rank, res))
    dist.destroy_process_group()


def test_allreduce_callback_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        assert out[r]["hist_ok"] and out[r]["band_hist_ok"]
        assert out[r]["sum"] == 4.0 and out[r]["min"] == 9.0 and out[r]["max"] == 11.0
        assert out[r]["rej"] == 201


RESHARD_CASES = [(8, 10, 6), (5, 7, 3), (1, 4, 4), (6, 1, 5)]


def _reshard_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from astroburst_b200.parallel import frame_owner, reshard_frames_to_bands, row_band

    ok = []
    for n_frames, rows, cols in RESHARD_CASES:
        full = torch.arange(n_frames * rows * cols, dtype=torch.float32).reshape(n_frames, rows, cols)
        mine = [f for f in range(n_frames) if frame_owner(f, n_frames, world) == rank]
        local = full[mine].contiguous()
        cube = reshard_frames_to_bands(local, n_frames, rows, cols)
        r0, r1 = row_band(rows, rank, world)
        good = bool(torch.equal(cube, full[:, r0:r1]))
        try:  # a shape that contradicts (rows, cols) is refused on every rank before any traffic
            reshard_frames_to_bands(local, n_frames, rows + 1, cols)
            good = False
        except ValueError:
            pass
        ok.append(good)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_reshard_frames_to_bands_gloo_world2():
    """Frame-sharded -> row-band sharded exchange (config 5): every rank ends with its row band of
    every frame; ragged frame counts / band sizes, a rank without frames and an empty band included."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_reshard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(out[0]) and all(out[1]), out


def test_band_pointers_in_cubes():
    """Pointer arithmetic of the peer-memory stack, checked on CPU tensors (data_ptr is plain)."""
    from astroburst_b200.parallel import band_pointers_in_cubes, frame_owner, row_band

    n_frames, rows, cols, world = 10, 7, 5, 3
    owners = [frame_owner(f, n_frames, world) for f in range(n_frames)]
    cubes = [torch.zeros((owners.count(r), rows, cols), dtype=torch.float32) for r in range(world)]
    bases = [c.data_ptr() for c in cubes]
    for rank in range(world):
        ptrs, (r0, r1) = band_pointers_in_cubes(bases, n_frames, rows, cols, rank, world)
        assert (r0, r1) == row_band(rows, rank, world) and len(ptrs) == n_frames
        seen = [0] * world
        for f, p in enumerate(ptrs):
            o = owners[f]
            if r1 > r0:
                assert p == cubes[o][seen[o], r0].data_ptr()
            seen[o] += 1


def test_c_abi_partition_rules_match_the_python_mirror():
    """ab_row_band / ab_frame_owner (shard.cu) == parallel.row_band / frame_owner (no GPU needed)."""
    import ctypes as C

    from astroburst_b200 import _lib
    from astroburst_b200.parallel import frame_owner, row_band

    lib = _lib.lib()
    for rows in (0, 1, 7, 64, 4096, 4099):
        for world in (1, 2, 3, 8, 16):
            for rank in range(world):
                r0, r1 = C.c_size_t(), C.c_size_t()
                lib.ab_row_band(rows, rank, world, C.byref(r0), C.byref(r1))
                assert (r0.value, r1.value) == row_band(rows, rank, world)
    for n in (1, 3, 8, 40, 255, 256):
        for world in (1, 2, 5, 8):
            assert [lib.ab_frame_owner(f, n, world) for f in range(n)] == [frame_owner(f, n, world) for f in range(n)]


def test_bind_to_gpu_numa_never_raises_and_keeps_a_usable_mask(monkeypatch):
    """Host placement is an optimisation: without NVML / a GPU it must leave the affinity mask alone
    and say so; with AB_NO_NUMA_BIND it must do nothing."""
    import os
    from astroburst_b200 import parallel

    before = os.sched_getaffinity(0)
    msg = parallel.bind_to_gpu_numa(0)
    assert isinstance(msg, str) and msg
    assert os.sched_getaffinity(0) and os.sched_getaffinity(0) <= before
    os.sched_setaffinity(0, before)
    monkeypatch.setenv("AB_NO_NUMA_BIND", "1")
    assert parallel.bind_to_gpu_numa(0).startswith("off")
    assert os.sched_getaffinity(0) == before
