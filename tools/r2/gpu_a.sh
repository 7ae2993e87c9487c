#!/bin/bash
# round 2, call A: new deep kernel -- parity, timing, bench line, ncu of the deep kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stack_gpu.py tests/test_calibration_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/r2/deep_time.py gpurun_out/r2_deep_timing.txt 2>&1 | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 3000 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stack_deep -s 2 -c 1 -o gpurun_out/r2_prof_deep -f python tools/deep_once.py 256 512 > gpurun_out/r2_prof_deep.log 2>&1; tail -3 gpurun_out/r2_prof_deep.log
