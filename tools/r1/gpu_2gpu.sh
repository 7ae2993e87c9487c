#!/bin/bash
# 2-GPU check: NCCL parity tests (sharded stats, band stack, frame-sharded stack) + the 2-GPU bench line.
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 3 2>gpurun_out/bench2.err | tee gpurun_out/bench_2gpu.json
python bench.py --steps 100 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench2.err | tee gpurun_out/bench_1gpu_samebox.json
tail -3 gpurun_out/bench2.err
