for mb in 96 32 192 384 768; do
  echo -n "band $mb MiB: "; AB_BAND_MB=$mb timeout 300 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline --e2e-steps 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(round(d['e2e']['value'],1))"
done
