#!/bin/bash
# Mid-round check: all GPU tests, probe A/B, per-pass stats timing, per-operator bench.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
bash tools/gpu_ab.sh
python tools/stats_once.py 2>&1 | tee gpurun_out/stats_once.txt
python bench_ops.py --out gpurun_out/bench_ops.json 2>&1 | grep -v '^{' | tail -5
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_ops.json"))
for x in d["results"]:
    print(f'{x["op"][:70]:70s} {x["ms"]:9.3f} ms {x["achieved_gbs"]:8.1f} GB/s frac {x["frac_of_measured_peak"]:.3f} parity {x.get("parity")}')
PY
