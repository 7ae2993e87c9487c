#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/r2_bench_${NG}gpu.json 2> gpurun_out/r2_bench_${NG}gpu.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_${NG}gpu.json") if l.startswith("{")][-1])
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "e2e", "multi_gpu_matches_oracle") if k in d})[:800])
    print("kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"])
    for k, v in d.get("secondary", {}).items():
        if isinstance(v, dict):
            print(k, {q: (round(x, 3) if isinstance(x, float) else x) for q, x in v.items() if q.startswith(("ms_", "speedup", "matches", "frac", "value", "max_corner"))})
        else:
            print(k, v)
except Exception as e:
    print("no bench line", e)
PY
tail -5 gpurun_out/r2_bench_${NG}gpu.err | cut -c1-300
