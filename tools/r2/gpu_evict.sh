for i in 1 2 3; do
  for v in default evict; do
    if [ $v = default ]; then unset AB_LIBRARY; else export AB_LIBRARY=$PWD/astroburst_b200/lib/variants/$v.so; fi
    echo -n "$v: "; timeout 300 python bench.py --steps 30 --warmup 5 --no-secondary 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(d['roofline']['kernel_ms'], d['roofline']['frac'], d['ms_per_step'])"
  done
done
