#!/bin/bash
# r2 final evidence on one GPU: tests, smoke, bench (both arms), per-operator bench, ncu launch list
# of the timed region + full capture of the headline kernel, stage traces.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/r2_bench_line_1gpu.json 2> gpurun_out/r2_bench_1gpu_err.txt; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference_line.json 2>/dev/null; echo "reference rc=$?"
timeout 900 python bench_ops.py > gpurun_out/bench_ops_r2.json 2> gpurun_out/bench_ops_err.txt; echo "bench_ops rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 20 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-secondary > gpurun_out/b_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:stack_ -s 3 -c 1 -f -o gpurun_out/r2_prof_stack python bench.py --steps 6 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-secondary > gpurun_out/prof.log 2>&1
AB_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --secondary c3 2> gpurun_out/r2_align_stage_trace_raw.txt > /dev/null
( echo "AB_TRACE=1 python bench.py --steps 5 --warmup 3 --secondary c3  (16 frames 8192^2; last frames of the last repetition)"; grep -E "detect 8192|star points|align 8192|match:" gpurun_out/r2_align_stage_trace_raw.txt | tail -14 ) > gpurun_out/r2_align_stage_trace.txt
timeout 300 python tools/r2/deep_time.py gpurun_out/r2_deep_timing.txt > /dev/null 2>&1
timeout 300 python tools/stats_once.py > gpurun_out/r2_stats_once.txt 2>&1
timeout 300 python tools/warp_once.py > gpurun_out/r2_warp_once.txt 2>&1
python tools/r2/c4_trace.py > gpurun_out/r2_c4_trace_1gpu.txt 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_line_1gpu.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'frac', d['roofline']['frac'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
for k,v in d['secondary'].items():
    if isinstance(v,dict): print(k, {a:b for a,b in v.items() if a in ('ms_total','value','speedup_vs_single_gpu','ms_per_target_frame_per_gpu','frac_of_roofline','matches_oracle','matches_oracle_on_sample','ms_single_gpu')})
PY
