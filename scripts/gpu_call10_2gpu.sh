#!/bin/bash
# 2-GPU debug: capture the errors of p2p_check / bench with remote-row staging.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== p2p_check"; timeout -k 10 240 $TR --master-port 29801 tests/mp/p2p_check.py > gpurun_out/p2p10.log 2>&1; echo "rc=$?"; grep -v "^\[W\|^W09\|OMP_NUM\|^\*\*\*" gpurun_out/p2p10.log | grep -B2 -A25 "Traceback\|Error\|error" | head -80; grep "ok:" gpurun_out/p2p10.log | tail -5
echo "== bench N=2"; timeout -k 10 240 $TR --master-port 29802 bench.py --gpus 2 --steps 20 --warmup 5 --no-arms --min-time 0.3 > gpurun_out/b10.json 2> gpurun_out/b10.err; echo "rc=$?"; cut -c1-600 gpurun_out/b10.json; grep -v "^\[W\|^W09\|OMP_NUM\|^\*\*\*\|NCCL" gpurun_out/b10.err | grep -B2 -A25 "Traceback\|Error\|error" | head -60
echo "== bench N=2 no staging"; GLT_B200_STAGE_REMOTE=0 timeout -k 10 240 $TR --master-port 29803 bench.py --gpus 2 --steps 20 --warmup 5 --no-arms --min-time 0.3 > gpurun_out/b10b.json 2> gpurun_out/b10b.err; echo "rc=$?"; cut -c1-300 gpurun_out/b10b.json; grep -v "^\[W\|^W09\|OMP_NUM\|^\*\*\*\|NCCL" gpurun_out/b10b.err | grep -B2 -A20 "Traceback\|Error\|error" | head -40
