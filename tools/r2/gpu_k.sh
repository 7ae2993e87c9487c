#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stack_gpu.py tests/test_calibration_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 300 python tools/r2/deep_time.py gpurun_out/r2_deep_timing_tma.txt 2>&1 | tail -8
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_stack_gpu.py -m gpu -x -q -k "deep_ragged or test_stack_matches_oracle" > gpurun_out/r2_racecheck_deep_tma.txt 2>&1; tail -4 gpurun_out/r2_racecheck_deep_tma.txt
