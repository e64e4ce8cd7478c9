#!/bin/bash
# GPU pass 23 (1 GPU): final check of the tree + ncu --set full of the default step (gather backward).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu23.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu23.log
echo "== smoke"; timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== ncu full step (default)"; timeout -k 10 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/step_full python bench.py --profile-steps 1 --no-pipeline > gpurun_out/ncu_step23.out 2>&1; echo "rc=$?"
python tools/ncu_summary.py /tmp/step_full.ncu-rep > gpurun_out/ncu_r2_step_gather_default_full_set.txt 2>/dev/null; wc -l gpurun_out/ncu_r2_step_gather_default_full_set.txt
ncu -i /tmp/step_full.ncu-rep --page source --csv -k regex:k_sage_gather_bwd 2>/dev/null | head -300 > gpurun_out/ncu_r2_gather_bwd_source_page_head.csv; wc -l gpurun_out/ncu_r2_gather_bwd_source_page_head.csv
echo "== bench default"; timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.7 2>/dev/null | grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'launches', d.get('gpu_launches'))"
