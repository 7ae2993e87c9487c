#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
AB_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --secondary c3 > gpurun_out/r2_p.json 2> gpurun_out/r2_p_trace.txt
grep -E "align 8192|match:" gpurun_out/r2_p_trace.txt | tail -4
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_p.json') if l.startswith('{')][-1])
c=d['secondary']['config3']; print(c['ms_total'], c['ms_per_target_frame_per_gpu'], c['matches_oracle'], c['max_corner_error_px'])"
