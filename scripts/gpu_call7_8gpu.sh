#!/bin/bash
# 8-GPU session: data-plane tests at 8 ranks, products scaling point with sections + in-kernel timeline,
# BASELINE config 3 (papers100M shape), config 4 (IGBH shape hetero), config 5 (SEAL on a 1 B-edge RMAT).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
flt() { grep -v "^\[W\|NCCL\|^W09\|^\*\*\*\|OMP_NUM\|^$" ; }
js() { grep '^{' | tail -1 ; }
echo "== p2p_check at $N ranks"; timeout -k 10 300 $TR --master-port 29601 tests/mp/p2p_check.py > gpurun_out/p2p_n$N.log 2>&1; echo "rc=$?"; grep "ok:\|ALL OK" gpurun_out/p2p_n$N.log | tail -4; grep -A8 Traceback gpurun_out/p2p_n$N.log | head -20
echo "== bench products N=$N"; timeout -k 10 300 $TR --master-port 29602 bench.py --gpus $N --steps 20 --warmup 5 --no-arms 2> gpurun_out/bench_products_n$N.err | js > gpurun_out/bench_products_n$N.json; echo "rc=$?"; cut -c1-1900 gpurun_out/bench_products_n$N.json; flt < gpurun_out/bench_products_n$N.err | tail -3
echo "== bench products N=$N staging off"; GLT_B200_STAGE_REMOTE=0 timeout -k 10 300 $TR --master-port 29610 bench.py --gpus $N --steps 20 --warmup 5 --no-arms --min-time 0.5 2>/dev/null | js | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('no-staging ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'l1', d['details']['layer1_autotune_ms'])"
echo "== bench products N=$N sections"; timeout -k 10 200 $TR --master-port 29603 bench.py --gpus $N --sections 2>&1 | grep sections_ms
echo "== bench products N=$N mxfp8"; timeout -k 10 300 $TR --master-port 29604 bench.py --gpus $N --steps 20 --warmup 5 --no-arms --min-time 0.5 --feat-format mxfp8 2>/dev/null | js | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('mxfp8 ms/step', round(d['ms_per_step'],4), 'value', round(d['value']))"
echo "== fused timeline N=$N"; GLT_B200_FUSED_TRACE=1 timeout -k 10 200 $TR --master-port 29605 tools/fused_trace.py --gpus $N 2>/dev/null | flt > gpurun_out/fused_trace_n$N.txt; head -12 gpurun_out/fused_trace_n$N.txt
echo "== papers100m N=$N"; timeout -k 10 600 $TR --master-port 29606 bench.py --gpus $N --shape papers100m --steps 20 --warmup 5 --no-arms 2> gpurun_out/bench_papers100m_n$N.err | js > gpurun_out/bench_papers100m_n$N.json; echo "rc=$?"; cut -c1-2300 gpurun_out/bench_papers100m_n$N.json; flt < gpurun_out/bench_papers100m_n$N.err | tail -4
echo "== hetero igbh-shape N=$N"; timeout -k 10 600 $TR --master-port 29607 benchmarks/bench_hetero_rgnn.py --papers 10000000 --feat-dim 1024 --hidden 512 --steps 20 2> gpurun_out/bench_hetero_n$N.err | js > gpurun_out/bench_hetero_n$N.json; echo "rc=$?"; cat gpurun_out/bench_hetero_n$N.json; flt < gpurun_out/bench_hetero_n$N.err | tail -4
echo "== seal 1B edges N=$N"; timeout -k 10 600 $TR --master-port 29608 benchmarks/bench_seal_subgraph.py --nodes 100000000 --edges 1000000000 --iters 20 2> gpurun_out/bench_seal_n$N.err | js > gpurun_out/bench_seal_n$N.json; echo "rc=$?"; cat gpurun_out/bench_seal_n$N.json; flt < gpurun_out/bench_seal_n$N.err | tail -4
