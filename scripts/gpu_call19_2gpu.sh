#!/bin/bash
# GPU pass 19 (2 GPUs): p2p_check (replicated topology, full replica, staged/in-place), pytest multi, bench N=2 with the
# placement policy (default) and fully partitioned (--replica-budget-gb 0), reference arm N=2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
one() { grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['ms_per_step'],4), 'l1', d['details'].get('layer1_autotune_ms'), d['details'].get('placement'), d['details'].get('hot_feature_replica'))"; }
echo "== p2p_check"; timeout -k 10 600 $TR --master-port 29511 tests/mp/p2p_check.py > gpurun_out/p2p_n2.log 2>&1; echo "rc=$?"; grep -E "ok:|ALL OK|Error|error|assert" gpurun_out/p2p_n2.log | grep "rank 0\|ALL OK\|rror" | cut -c1-200 | tail -24
echo "== pytest multi"; timeout -k 10 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -3
echo "== bench N=2 default placement"; timeout -k 10 600 $TR --master-port 29512 bench.py --gpus $N --no-arms 2>gpurun_out/b2a.err | tee gpurun_out/bench_r2_final_2gpu.json | one
echo "== bench N=2 partitioned"; timeout -k 10 600 $TR --master-port 29513 bench.py --gpus $N --no-arms --replica-budget-gb 0 2>gpurun_out/b2b.err | tee gpurun_out/bench_r2_final_2gpu_partitioned.json | one
echo "== reference N=2"; timeout -k 10 900 $TR --master-port 29514 bench.py --impl reference --gpus $N 2>gpurun_out/b2r.err | grep '^{' | tail -1 | tee gpurun_out/bench_r2_final_reference_2gpu.json | cut -c1-400
