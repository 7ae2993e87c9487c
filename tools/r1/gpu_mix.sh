#!/bin/bash
# Stack tests, then one short bench line per compare-exchange mix, then an ncu capture of the default.
mkdir -p gpurun_out
python -m pytest tests/test_stack_gpu.py -m gpu -x -q 2>&1 | tail -5
for m in ${MIXES:-0 1 2 3 4 5 6 7}; do tools/bench_ms.sh AB_STACK_MIX=$m; done 2>&1 | tee gpurun_out/mix.txt
ncu --set full --clock-control none --import-source on -k regex:stack_ -s 3 -c 1 -f -o gpurun_out/prof_stack python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/prof.log 2>&1
ls -la gpurun_out | head -5
