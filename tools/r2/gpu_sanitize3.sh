#!/bin/bash
mkdir -p gpurun_out
( echo "compute-sanitizer --tool memcheck: tests/test_fits_gpu.py (ab_stack_fits_units), tests/test_stack_gpu.py -k frame_counts_between (padded register route), tests/test_pipelines_gpu.py -k align (from-stars entry), tests/test_stars_gpu.py -k 'tie_heavy or device_ransac'"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_fits_gpu.py -m gpu -q -x 2>&1 | tail -4
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_stack_gpu.py -m gpu -q -x -k "frame_counts_between" 2>&1 | tail -4
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_pipelines_gpu.py -m gpu -q -x -k "align" 2>&1 | tail -4
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_stars_gpu.py -m gpu -q -x -k "tie_heavy or device_ransac" 2>&1 | tail -4 ) > gpurun_out/r2_memcheck_late.txt 2>&1
cat gpurun_out/r2_memcheck_late.txt
