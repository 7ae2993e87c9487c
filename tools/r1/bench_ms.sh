#!/bin/bash
# usage: tools/bench_ms.sh [ENV=VAL ...]  -> prints ms_per_step / Mpix/s / frac of a short bench run
env "$@" python bench.py --steps 20 --warmup 3 --e2e-steps 1 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$*', 'ms_per_step=%.4f'%d['ms_per_step'], 'Mpix/s=%.0f'%d['value'], 'frac=%.3f'%d['roofline']['frac'], 'rej=',d['config']['rejected_pixels'], 'clk=',d['clocks']['sm_mhz'])"
