#!/bin/bash
mkdir -p gpurun_out
python tools/r2/deep_ab.py tma 2>&1 | tail -1
AB_LIBRARY=$PWD/astroburst_b200/lib/variants/deep_notma.so python tools/r2/deep_ab.py ldgsts 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stack_deep -s 1 -c 1 -o gpurun_out/r2_prof_deep_tma -f python tools/deep_once.py 256 512 > gpurun_out/r2_prof_deep_tma.log 2>&1; tail -2 gpurun_out/r2_prof_deep_tma.log
