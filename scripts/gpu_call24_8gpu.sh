#!/bin/bash
# GPU pass 24 (8 GPUs): the final tree at N=8 (one-part full replica, adam_peer load order).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=8
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --no-arms 2>gpurun_out/b8f.err | grep '^{' | tail -1 > gpurun_out/bench_r2_final_8gpu_v2.json
python -c "import json; d=json.load(open('gpurun_out/bench_r2_final_8gpu_v2.json')); print('ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['ms_per_step'],4), 'l1', d['details'].get('layer1_autotune_ms'), d['details'].get('placement'), d['details'].get('hot_feature_replica'), d['clocks'])" || tail -5 gpurun_out/b8f.err
