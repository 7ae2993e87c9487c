#!/bin/bash
# 2-GPU call: multi-GPU parity tests + bench at N=2
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -25
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_2gpu.json") if l.startswith("{")][-1])
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "roofline", "e2e", "multi_gpu_matches_oracle") if k in d})[:1500])
    print(json.dumps(d.get("secondary"), indent=1)[:7000])
except Exception as e:
    print("no bench line", e)
PY
tail -30 gpurun_out/r2_bench_2gpu.err
