#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_deep_ab.txt
python tools/r2/deep_ab.py default 2>&1 | tail -1 | tee -a gpurun_out/r2_deep_ab.txt
for v in deep_x deep_m11 deep_m34 deep_m12 deep_x_m11; do
  AB_LIBRARY=$PWD/astroburst_b200/lib/variants/$v.so timeout 300 python tools/r2/deep_ab.py $v 2>&1 | tail -1 | tee -a gpurun_out/r2_deep_ab.txt
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
