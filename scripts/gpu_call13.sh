#!/bin/bash
# GPU pass 13 (1 GPU): in-situ per-launch times; merged zero launch; PDL on the training stream only.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
one() { grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'launches', d.get('gpu_launches'), 'loss', d['details'].get('last_loss'))"; }
echo "== engine tests"; timeout -k 10 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x 2>&1 | tail -3
echo "== kernel times"; timeout -k 10 300 python bench.py --kernel-times 2>&1 | grep -v Warning | tee gpurun_out/kernel_times_insitu.txt
B="python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.7"
echo "== bench default";            timeout -k 10 300 $B 2>/dev/null | one
echo "== bench MERGED_ZERO=1";      GLT_B200_MERGED_ZERO=1 timeout -k 10 300 $B 2>/dev/null | one
echo "== bench PDL_MODE=train";     GLT_B200_PDL_MODE=train timeout -k 10 300 $B 2>/dev/null | one
echo "== bench MERGED_ZERO+PDL train"; GLT_B200_MERGED_ZERO=1 GLT_B200_PDL_MODE=train timeout -k 10 300 $B 2>/dev/null | one
echo "== bench no-pipeline (PDL all, single stream)"; timeout -k 10 300 $B --no-pipeline 2>/dev/null | one
