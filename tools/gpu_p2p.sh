#!/bin/bash
# 2-GPU: NCCL parity tests incl. the peer-memory (P2P) frame-sharded stack, then config 5 at N=2 both ways.
mkdir -p gpurun_out
python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -15
for mode in "" "--p2p"; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench_config5.py --frames 256 --rows 1024 --steps 4 $mode 2>gpurun_out/c5.err | grep '^{' | tee -a gpurun_out/config5_2gpu.json
tail -2 gpurun_out/c5.err | cut -c1-300
done
