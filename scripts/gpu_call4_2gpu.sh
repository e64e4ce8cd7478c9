#!/bin/bash
# 2-GPU validation: NVLink data plane tests, multi-GPU bench build path, hetero engine with peer all-reduce.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== multi-gpu pytest"; timeout -k 10 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_multi.log
echo "== bench N=2"; timeout -k 10 400 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-arms --min-time 0.5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; cat gpurun_out/bench_n2.json; grep -v "^\[W\|NCCL\|^$" gpurun_out/bench_n2.err | tail -5
echo "== bench N=2 sections"; timeout -k 10 300 $TR --master-port 29512 bench.py --gpus 2 --sections 2>&1 | grep sections_ms
echo "== bench papers100m-shape (scaled 1/20) N=2"; timeout -k 10 400 $TR --master-port 29513 bench.py --gpus 2 --shape papers100m --nodes 5500000 --edges 80000000 --steps 20 --warmup 5 --no-arms --min-time 0.3 > gpurun_out/bench_p100m_small_n2.json 2> gpurun_out/bench_p100m_small_n2.err; echo "rc=$?"; cat gpurun_out/bench_p100m_small_n2.json; grep -v "^\[W\|NCCL\|^$" gpurun_out/bench_p100m_small_n2.err | tail -5
echo "== hetero engine N=2"; timeout -k 10 400 $TR --master-port 29514 benchmarks/bench_hetero_rgnn.py --papers 200000 --feat-dim 256 --hidden 256 2>&1 | grep -v "^\[W\|NCCL" | tail -4
echo "== reference N=2"; timeout -k 10 600 $TR --master-port 29515 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 --no-arms --min-time 0.5 2>/dev/null | tail -1
