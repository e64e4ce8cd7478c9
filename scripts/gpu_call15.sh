#!/bin/bash
# GPU pass 15 (1 GPU): rewritten gather backward (32-row chunks, index chain per lane, 4-warp teams, red.v4 column sums).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
one() { grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'launches', d.get('gpu_launches'), 'loss', d['details'].get('last_loss'))"; }
echo "== engine tests"; timeout -k 10 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x 2>&1 | tail -4
echo "== kernel times gather-bwd"; GLT_B200_GATHER_BWD=1 timeout -k 10 300 python bench.py --kernel-times 2>&1 | grep -v Warning | tee gpurun_out/kernel_times_insitu_gatherbwd_v2.txt
B="python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.7"
echo "== bench default";       timeout -k 10 300 $B 2>/dev/null | one
echo "== bench GATHER_BWD=1";  GLT_B200_GATHER_BWD=1 timeout -k 10 300 $B 2>/dev/null | one
echo "== bench GATHER_BWD=1 dropout 0.5";  GLT_B200_GATHER_BWD=1 timeout -k 10 300 $B --dropout 0.5 2>/dev/null | one
