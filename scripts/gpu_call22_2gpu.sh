#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== kernel times N=2 (replicated data)"; timeout -k 10 400 $TR --master-port 29531 bench.py --gpus $N --kernel-times 2>&1 | grep -E " us  |^#" | tee gpurun_out/kernel_times_insitu_2gpu.txt
