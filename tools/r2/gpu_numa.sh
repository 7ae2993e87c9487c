nproc; lscpu | grep -i "numa\|socket\|model name" | head -8; nvidia-smi topo -m 2>/dev/null | head -14
for v in bind nobind bind nobind; do
  if [ $v = nobind ]; then export AB_NO_NUMA_BIND=1; else unset AB_NO_NUMA_BIND; fi
  echo -n "$v: "; timeout 300 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(d['e2e'])"
done
