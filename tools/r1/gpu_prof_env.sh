#!/bin/bash
# usage: tools/gpu_prof_env.sh OUTNAME [ENV=VAL ...] : ncu full capture of one stack kernel launch
out=$1; shift
mkdir -p gpurun_out
timeout 200 env "$@" ncu --set full --clock-control none --import-source on -k regex:stack_ -s 3 -c 1 -f -o gpurun_out/$out python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/$out.log 2>&1
ls -la gpurun_out/$out.ncu-rep
