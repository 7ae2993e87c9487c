#!/bin/bash
# ncu full capture of the stack kernel only (one launch), plus a short bench line.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --e2e-steps 2 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench.json
ncu --set full --clock-control none --import-source on -k regex:stack_ -s 3 -c 1 -f -o gpurun_out/prof_stack python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/prof.log 2>&1
ls -la gpurun_out
