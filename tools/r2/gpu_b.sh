#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stack_deep -s 1 -c 1 -o gpurun_out/r2_prof_deep -f python tools/deep_once.py 256 512 > gpurun_out/r2_prof_deep.log 2>&1; tail -3 gpurun_out/r2_prof_deep.log
