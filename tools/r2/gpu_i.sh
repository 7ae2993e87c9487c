#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench_configs.py --small 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c4 small ms', d['config4']['ms_total'], d['config4']['matches_oracle'], '| c3 small ms', d['config3']['ms_total'], d['config3']['matches_oracle'])"
timeout 600 python bench.py --steps 20 --warmup 5 --secondary c4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c4 full', d['secondary']['config4']['ms_total'], d['secondary']['config4']['matches_oracle'])"
