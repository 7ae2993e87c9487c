#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipelines_gpu.py tests/test_imaging_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench_configs.py --small 2>&1 | tail -3
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err ) 2>&1 | tail -4
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_f.json"))
print(json.dumps(d.get("secondary"), indent=1)[:6000])
PY
tail -5 gpurun_out/r2_bench_f.err
