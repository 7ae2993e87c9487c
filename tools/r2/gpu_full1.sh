mkdir -p gpurun_out
S=$(date +%s)
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "pytest wall $(( $(date +%s) - S )) s"
S=$(date +%s)
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "smoke wall $(( $(date +%s) - S )) s"
S=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r2_bench_line_1gpu.json 2> gpurun_out/r2_bench_1gpu_err.txt
echo "bench wall $(( $(date +%s) - S )) s"; tail -2 gpurun_out/r2_bench_1gpu_err.txt
S=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference_line.json 2>/dev/null
echo "reference arm wall $(( $(date +%s) - S )) s"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_line_1gpu.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'])
for k,v in d['secondary'].items():
    if isinstance(v,dict): print(k, {a:b for a,b in v.items() if a in ('ms_total','value','speedup_vs_single_gpu','ms_per_target_frame_per_gpu','frac_of_roofline','matches_oracle','matches_oracle_on_sample','ms_stack','ms_single_gpu')})
    else: print(k,v)
r=json.loads([l for l in open('gpurun_out/r2_bench_reference_line.json') if l.startswith('{')][-1])
print('reference', r['value'], r['cpu_baseline'])
PY
