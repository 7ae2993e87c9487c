#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench_ops.py --out gpurun_out/r2_bench_ops.json 2>/dev/null | grep -E "compute_image_stats|colour chain" | cut -c1-260
AB_STATS_HOST_WALKS=1 python bench_ops.py --out gpurun_out/r2_bench_ops_hostwalk.json 2>/dev/null | grep -E "compute_image_stats" | cut -c1-260
