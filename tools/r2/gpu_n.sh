#!/bin/bash
mkdir -p gpurun_out
AB_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --secondary c3 > gpurun_out/r2_n.json 2> gpurun_out/r2_n_trace.txt
grep "align 8192" gpurun_out/r2_n_trace.txt | tail -3
grep "detect 8192" gpurun_out/r2_n_trace.txt | tail -2
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_n.json') if l.startswith('{')][-1])
print(d['secondary']['config3']['ms_total'], d['secondary']['config3']['ms_per_target_frame_per_gpu'])"
