#!/bin/bash
mkdir -p gpurun_out
python tools/r2/arcsinh_err.py 2>&1 | tail -6
timeout 600 python -m pytest tests/test_cpp_header_gpu.py -m gpu -x -q 2>&1 | tail -4
# racecheck / memcheck of the stack kernels (TMA stage re-arm without a proxy fence; the new deep kernel)
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_stack_gpu.py -m gpu -x -q -k "stack_matches_oracle or deep_ragged or reference_unit" > gpurun_out/r2_racecheck_stack.txt 2>&1; tail -12 gpurun_out/r2_racecheck_stack.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_stack_gpu.py -m gpu -x -q -k "stack_matches_oracle or deep_ragged or deep_far or single_pixel_random" > gpurun_out/r2_memcheck_stack.txt 2>&1; tail -8 gpurun_out/r2_memcheck_stack.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_stars_gpu.py tests/test_calibration_gpu.py -m gpu -x -q -k "components_on_device or mean_stack or subframe" > gpurun_out/r2_memcheck_stars.txt 2>&1; tail -8 gpurun_out/r2_memcheck_stars.txt
