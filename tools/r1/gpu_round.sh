#!/bin/bash
# One GPU call: tests, smoke, bench, ncu launch list + full capture of the stack kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py 2>gpurun_out/bench.err | tee gpurun_out/bench.json
python bench.py --impl reference --steps 5 --warmup 1 | tee gpurun_out/bench_ref.json
# launch list of the timed region only (bench.py brackets it with cudaProfilerStart/Stop)
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:stack_ -s 3 -c 1 -f -o gpurun_out/prof_stack python bench.py --steps 6 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/prof.log 2>&1
ls -la gpurun_out
python bench_ops.py --out gpurun_out/bench_ops.json 2>&1 | tail -20
