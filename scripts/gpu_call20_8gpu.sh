#!/bin/bash
# GPU pass 20 (8 GPUs): final scaling numbers -- products (placement policy / fully partitioned), papers100m.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
one() { grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['ms_per_step'],4), 'l1', d['details'].get('layer1_autotune_ms'), d['details'].get('placement'), d['details'].get('hot_feature_replica'))"; }
echo "== bench products N=$N default placement"; timeout -k 10 400 $TR --master-port 29521 bench.py --gpus $N --no-arms 2>gpurun_out/b8a.err | tee gpurun_out/bench_r2_final_${N}gpu.json | one
echo "== bench products N=$N partitioned"; timeout -k 10 400 $TR --master-port 29522 bench.py --gpus $N --no-arms --replica-budget-gb 0 2>gpurun_out/b8b.err | tee gpurun_out/bench_r2_final_${N}gpu_partitioned.json | one
echo "== sections N=$N default placement"; timeout -k 10 300 $TR --master-port 29523 bench.py --gpus $N --sections 2>/dev/null | grep '^{' | tail -1
echo "== papers100m N=$N default placement"; timeout -k 10 700 $TR --master-port 29524 bench.py --gpus $N --shape papers100m --no-arms 2>gpurun_out/b8p.err | tee gpurun_out/bench_r2_final_papers100m_${N}gpu.json | one
