#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
one() { grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'launches', d.get('gpu_launches'), 'loss', d['details'].get('last_loss'))"; }
echo "== gather tests"; timeout -k 10 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -k "gather or dropout or transposed" 2>&1 | grep -E "^E       Assert|^E       assert|^FAILED|passed|failed" | cut -c1-200 | head -12
echo "== kernel times gather-bwd"; GLT_B200_GATHER_BWD=1 timeout -k 10 300 python bench.py --kernel-times 2>&1 | grep -v Warning | tee gpurun_out/kernel_times_insitu_gatherbwd_v3.txt
B="python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.7"
echo "== bench GATHER_BWD=1";  GLT_B200_GATHER_BWD=1 timeout -k 10 300 $B 2>/dev/null | one
