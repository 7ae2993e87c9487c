mkdir -p gpurun_out
( nproc; lscpu | grep -i "numa\|socket"; nvidia-smi topo -m | head -12 ) > gpurun_out/r2_8gpu_topology.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -3
for N in 8 2 4; do
  timeout 900 $TR --nproc-per-node $N --master-port $((29500+N)) bench.py --gpus $N > gpurun_out/r2_bench_line_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu_err.txt
  echo "N=$N rc=$?"; tail -1 gpurun_out/r2_bench_${N}gpu_err.txt | cut -c1-200
done
AB_NO_NUMA_BIND=1 timeout 600 $TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --steps 10 --no-secondary --no-cpu-baseline > gpurun_out/r2_bench_line_8gpu_nobind.json 2>/dev/null
timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 10 --no-secondary --no-cpu-baseline > gpurun_out/r2_bench_line_8gpu_bind.json 2>/dev/null
python - <<'PY'
import json
def last(p):
    try: return json.loads([l for l in open(p) if l.startswith('{')][-1])
    except Exception as e: return None
for n in (2,4,8):
    d=last(f'gpurun_out/r2_bench_line_{n}gpu.json')
    if not d: print(n,'no line'); continue
    print(n, round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', round(d['e2e']['value']), d['e2e'].get('host_placement'))
    for k,v in d['secondary'].items():
        if isinstance(v,dict): print('  ',k, {a:b for a,b in v.items() if a in ('ms_total','value','speedup_vs_single_gpu','ms_per_target_frame_per_gpu','matches_oracle','matches_oracle_on_sample','ms_stack','ms_single_gpu','ms_band_resident')})
        else: print('  ',k,v)
for t in ('nobind','bind'):
    d=last(f'gpurun_out/r2_bench_line_8gpu_{t}.json')
    print(t, d and (round(d['e2e']['value']), d['e2e'].get('host_placement')))
PY
cat gpurun_out/r2_8gpu_topology.txt | head -20
