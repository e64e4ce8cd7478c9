#!/bin/bash
# 2-GPU: remote-row staging validation + A/B (second attempt, explicit remote marking).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
one() { grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'fused', d['details']['fused_tcgen05_layer1'], 'l1', d['details']['layer1_autotune_ms'], 'staged', d['details'].get('remote_rows_staged_on_sampling_stream'))"; }
echo "== p2p_check"; timeout -k 10 300 $TR --master-port 29801 tests/mp/p2p_check.py > gpurun_out/p2p11.log 2>&1; echo "rc=$?"; grep -v "^\[W\|^W09\|OMP_NUM\|^\*\*\*" gpurun_out/p2p11.log | grep -A12 "Traceback" | head -40; grep "ok:\|ALL OK" gpurun_out/p2p11.log | tail -12
echo "== bench N=2 staging on"; timeout -k 10 300 $TR --master-port 29802 bench.py --gpus 2 --steps 20 --warmup 5 --no-arms --min-time 0.5 2>gpurun_out/b11a.err | one
echo "== bench N=2 staging off"; GLT_B200_STAGE_REMOTE=0 timeout -k 10 300 $TR --master-port 29803 bench.py --gpus 2 --steps 20 --warmup 5 --no-arms --min-time 0.5 2>gpurun_out/b11b.err | one
echo "== bench N=2 staging on, fused forced"; timeout -k 10 300 $TR --master-port 29804 bench.py --gpus 2 --steps 20 --warmup 5 --no-arms --min-time 0.5 --fused on 2>/dev/null | one
echo "== bench N=2 mxfp8"; timeout -k 10 300 $TR --master-port 29805 bench.py --gpus 2 --steps 20 --warmup 5 --no-arms --min-time 0.5 --feat-format mxfp8 2>/dev/null | one
echo "== sections N=2 staging on"; timeout -k 10 300 $TR --master-port 29806 bench.py --gpus 2 --sections --fused on 2>&1 | grep sections_ms
