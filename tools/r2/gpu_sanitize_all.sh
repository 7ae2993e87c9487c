#!/bin/bash
mkdir -p gpurun_out
( echo "compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q   (whole GPU suite, final r2 tree)"
  timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r2_memcheck_all.txt 2>&1
tail -6 gpurun_out/r2_memcheck_all.txt
( echo "compute-sanitizer --tool racecheck python -m pytest tests/test_imaging_gpu.py tests/test_fits_gpu.py tests/test_calibration_gpu.py -m gpu -q"
  timeout 2400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_imaging_gpu.py tests/test_fits_gpu.py tests/test_calibration_gpu.py -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r2_racecheck_imaging.txt 2>&1
tail -5 gpurun_out/r2_racecheck_imaging.txt
