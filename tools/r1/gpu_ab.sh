#!/bin/bash
# A/B: short bench lines for the default library and every lib/variants/*.so, two interleaved rounds.
mkdir -p gpurun_out
for round in 1 2; do
  tools/bench_ms.sh AB_STACK_MIX=${MIX:-0}
  for so in astroburst_b200/lib/variants/*.so; do tools/bench_ms.sh AB_LIBRARY=$PWD/$so AB_STACK_MIX=${MIX:-0}; done
done 2>&1 | sed -e "s#$PWD/astroburst_b200/lib/variants/##" | tee gpurun_out/ab.txt
