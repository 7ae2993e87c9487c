#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_stack_gpu.py tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -6
for tma in 1 0; do
AB_DEEP_FORCE_TMA=$tma timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $NG --steps 10 --warmup 3 --secondary c5 > gpurun_out/r2_m_tma$tma.json 2> gpurun_out/r2_m_tma$tma.err
python - $tma <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2_m_tma{tag}.json") if l.startswith("{")][-1])
    s = d.get("secondary", {})
    print("tma", tag, {q: (round(v, 3) if isinstance(v, float) else v) for q, v in s.get("config5", {}).items() if q.startswith(("ms_", "speedup", "matches", "frac"))}, s.get("error"))
except Exception as e:
    print(tag, "no line", e)
PY
done
