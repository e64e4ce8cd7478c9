#!/bin/bash
# GPU pass 12 (1 GPU): full GPU test-suite (dropout tests new), A/B of the aggregate batching and side-stream zero fill.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
one() { grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'launches', d.get('gpu_launches'), 'loss', d['details'].get('last_loss'))"; }
echo "== pytest -m gpu"; timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu12.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu12.log
B="python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.7"
echo "== bench default";            timeout -k 10 300 $B 2>/dev/null | one
echo "== bench AGG_BATCH=1";        GLT_B200_AGG_BATCH=1 timeout -k 10 300 $B 2>/dev/null | one
echo "== bench SIDE_ZERO=1";        GLT_B200_SIDE_ZERO=1 timeout -k 10 300 $B 2>/dev/null | one
echo "== bench AGG_BATCH+SIDE_ZERO"; GLT_B200_AGG_BATCH=1 GLT_B200_SIDE_ZERO=1 timeout -k 10 300 $B 2>/dev/null | one
echo "== bench dropout 0.5";        timeout -k 10 300 $B --dropout 0.5 2>/dev/null | one
echo "== bench default again";      timeout -k 10 300 $B 2>/dev/null | one
