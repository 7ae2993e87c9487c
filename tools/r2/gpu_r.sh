#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stack_gpu.py tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -4
run() { env "${@:2}" timeout 300 python bench.py --steps 100 --warmup 10 --no-secondary --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', 'ms_per_step', round(d['ms_per_step'],5), 'kernel_ms', round(d['roofline']['kernel_ms'],5), 'frac', round(d['roofline']['frac'],4), 'rej', d['result']['rejected_pixels_all_ranks'])"; }
run default AB_X=1
run default AB_X=1
for v in hl_s1516 hl_s34 hl_s1; do run $v AB_LIBRARY=$PWD/astroburst_b200/lib/variants/$v.so; done
