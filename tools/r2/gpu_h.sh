#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -25
run() { # $1 = tag, rest = env
  timeout 600 env "${@:2}" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --secondary $SEC > gpurun_out/r2_h_$1.json 2> gpurun_out/r2_h_$1.err
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2_h_{tag}.json") if l.startswith("{")][-1])
    s = d.get("secondary", {})
    for k in ("config5", "config4"):
        if k in s:
            print(tag, k, {q: (round(v, 3) if isinstance(v, float) else v) for q, v in s[k].items() if q.startswith(("ms_", "speedup", "matches", "frac"))})
    if "error" in s: print(tag, "ERROR", s["error"])
except Exception as e:
    print(tag, "no line", e)
PY
  tail -3 gpurun_out/r2_h_$1.err | cut -c1-300
}
SEC=c5,c4 run default AB_X=1
SEC=c5 run ca AB_LIBRARY=$PWD/astroburst_b200/lib/variants/deep_ca.so
