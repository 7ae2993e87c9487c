#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
AB_TRACE=1 python bench_ops.py --out gpurun_out/bench_ops.json 2>gpurun_out/trace.txt | grep -E "detect_stars|align_channel|masked_stretch|warp_image|exact select" | cut -c1-230
sort gpurun_out/trace.txt | uniq -c | sort -rn | head -8
