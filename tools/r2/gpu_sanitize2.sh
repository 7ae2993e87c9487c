#!/bin/bash
# compute-sanitizer over the kernels added after the first r2 sanitizer pass
mkdir -p gpurun_out
SEL='background or detect_matches or bucketed or ransac or match_and_ransac or align_recovers'
( echo "compute-sanitizer --tool memcheck python -m pytest tests/test_stars_gpu.py -k \"$SEL\" + tests/test_imaging_gpu.py -k \"bicubic or warp\""
  timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_stars_gpu.py -m gpu -q -x -k "$SEL" 2>&1 | tail -8
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_imaging_gpu.py -m gpu -q -x -k "bicubic or warp" 2>&1 | tail -8 ) > gpurun_out/r2_memcheck_stars2_resample.txt
( echo "compute-sanitizer --tool racecheck python -m pytest tests/test_stars_gpu.py -k \"background or bucketed or ransac\""
  timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_stars_gpu.py -m gpu -q -x -k "background or bucketed or ransac" 2>&1 | tail -12 ) > gpurun_out/r2_racecheck_stars2.txt
tail -5 gpurun_out/r2_memcheck_stars2_resample.txt gpurun_out/r2_racecheck_stars2.txt
