#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stars_gpu.py tests/test_golden_gpu.py tests/test_preview_gpu.py -m gpu -x -q 2>&1 | tail -15
: > gpurun_out/r2_deep_ab2.txt
python tools/r2/deep_ab.py default 2>&1 | tail -1 | tee -a gpurun_out/r2_deep_ab2.txt
for v in deep_w6 deep_w12; do
  AB_LIBRARY=$PWD/astroburst_b200/lib/variants/$v.so timeout 300 python tools/r2/deep_ab.py $v 2>&1 | tail -1 | tee -a gpurun_out/r2_deep_ab2.txt
done
AB_TRACE=1 timeout 600 python bench_ops.py --out gpurun_out/r2_bench_ops.json 2>gpurun_out/r2_trace.txt | grep -E "detect_stars|align_channel|warp|stats" | cut -c1-220
grep "detect 8192" gpurun_out/r2_trace.txt | tail -3
