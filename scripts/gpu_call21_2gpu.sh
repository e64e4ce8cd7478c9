#!/bin/bash
# GPU pass 21 (2 GPUs): single-part full feature replica.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
one() { grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['ms_per_step'],4), 'l1', d['details'].get('layer1_autotune_ms'), d['details'].get('placement'), d['details'].get('hot_feature_replica'))"; }
echo "== p2p_check"; timeout -k 10 600 $TR --master-port 29511 tests/mp/p2p_check.py > gpurun_out/p2p_n2.log 2>&1; echo "rc=$?"; grep -E "ok:|ALL OK|Error|error|assert" gpurun_out/p2p_n2.log | grep "rank 0\|ALL OK\|rror" | cut -c1-160 | tail -20
echo "== bench N=2 default placement"; timeout -k 10 600 $TR --master-port 29512 bench.py --gpus $N --no-arms 2>gpurun_out/b2a.err | tee gpurun_out/bench_r2_final_2gpu.json | one
tail -3 gpurun_out/b2a.err
