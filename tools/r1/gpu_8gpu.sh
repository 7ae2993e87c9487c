#!/bin/bash
# 8-GPU check: the headline bench at N=8 (weak scaling) and BASELINE config 5 (256 frames sharded over 8 GPUs).
mkdir -p gpurun_out
nvidia-smi -L | wc -l
N=${N:-8}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 100 --warmup 3 --e2e-steps 2 2>gpurun_out/bench8.err | tee gpurun_out/bench_${N}gpu.json | cut -c1-400
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench_config5.py --steps 4 2>>gpurun_out/bench8.err | tee gpurun_out/config5_${N}gpu.json
tail -3 gpurun_out/bench8.err | cut -c1-300
