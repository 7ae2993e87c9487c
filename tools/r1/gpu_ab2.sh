#!/bin/bash
mkdir -p gpurun_out
for round in 1 2; do
  tools/bench_ms.sh AB_STACK_MIX=0
  tools/bench_ms.sh AB_STACK_MIX=5
  tools/bench_ms.sh AB_STACK_MIX=0 AB_STACK_VARIANT=2
  tools/bench_ms.sh AB_STACK_MIX=0 AB_STACK_VARIANT=2 AB_LIBRARY=$PWD/astroburst_b200/lib/variants/ldg4.so
  tools/bench_ms.sh AB_STACK_MIX=5 AB_STACK_VARIANT=2 AB_LIBRARY=$PWD/astroburst_b200/lib/variants/ldg4.so
done 2>&1 | sed -e "s#$PWD/astroburst_b200/lib/variants/##" | tee gpurun_out/ab2.txt
