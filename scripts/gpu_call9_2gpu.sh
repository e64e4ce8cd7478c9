#!/bin/bash
# 2-GPU: remote-row staging validation + A/B, MXFP8 features, NVLink counters of the in-kernel peer loads.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
flt() { grep -v "^\[W\|NCCL\|^W09\|^\*\*\*\|OMP_NUM\|^$" ; }
one() { python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'fused', d['details']['fused_tcgen05_layer1'], 'l1', d['details']['layer1_autotune_ms'], 'staged', d['details'].get('remote_rows_staged_on_sampling_stream'))"; }
echo "== p2p_check"; timeout -k 10 600 $TR --master-port 29701 tests/mp/p2p_check.py 2>&1 | flt | tail -14
echo "== bench N=2 staging on"; timeout -k 10 400 $TR --master-port 29702 bench.py --gpus 2 --steps 20 --warmup 5 --no-arms --min-time 0.5 2>/dev/null | one
echo "== bench N=2 staging off"; GLT_B200_STAGE_REMOTE=0 timeout -k 10 400 $TR --master-port 29703 bench.py --gpus 2 --steps 20 --warmup 5 --no-arms --min-time 0.5 2>/dev/null | one
echo "== bench N=2 staging on, fused forced"; timeout -k 10 400 $TR --master-port 29704 bench.py --gpus 2 --steps 20 --warmup 5 --no-arms --min-time 0.5 --fused on 2>/dev/null | one
echo "== bench N=2 mxfp8"; timeout -k 10 400 $TR --master-port 29705 bench.py --gpus 2 --steps 20 --warmup 5 --no-arms --min-time 0.5 --feat-format mxfp8 2>/dev/null | one
echo "== sections N=2"; timeout -k 10 300 $TR --master-port 29706 bench.py --gpus 2 --sections 2>&1 | grep sections_ms
echo "== nvlink counters"; timeout -k 10 700 bash tools/ncu_peer.sh 2>&1 | tail -30
cat gpurun_out/nvlink_metric_names.txt
