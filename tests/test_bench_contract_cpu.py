"""bench.py's reference arm runs on the host alone (`--impl reference`): it must print one JSON line
with the contract's keys, and the non-zero ranks of a torchrun launch must exit without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("Mpix/s sigma-clip stack") and d["value"] > 0 and d["steps"] == 1
    assert d["config"]["workload"].startswith("32-frame sigma-clip stack, 4096x4096 f32")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None


def test_reference_arm_other_ranks_do_nothing():
    r = _run({"RANK": "3", "WORLD_SIZE": "8", "LOCAL_RANK": "3"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_uses_all_host_threads_under_torchrun_env():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm must still use every host core (r1's
    multi-GPU vs_reference was void because it ran on one thread) and carry the GPU arm's config."""
    sys.path.insert(0, ROOT)
    import bench

    r = _run({"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "8", "LOCAL_RANK": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["config"] == bench.CONFIG
