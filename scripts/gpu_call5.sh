#!/bin/bash
# GPU pass 5: PDL on/off, fp8 features, full suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full gpu suite (PDL on)"; timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu5.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu5.log
one() { python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'k/step', d['kernels_per_step'], 'l1', d['details']['layer1_autotune_ms'], 'loss', round(d['details']['last_loss'],3))"; }
echo "== bench PDL on"; timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.5 2>gpurun_out/b5a.err | one; tail -2 gpurun_out/b5a.err
echo "== bench PDL off"; GLT_B200_PDL=0 timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.5 2>gpurun_out/b5b.err | one; tail -2 gpurun_out/b5b.err
echo "== bench PDL on, no pipeline"; timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.5 --no-pipeline 2>/dev/null | one
echo "== bench PDL off, no pipeline"; GLT_B200_PDL=0 timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.5 --no-pipeline 2>/dev/null | one
echo "== bench mxfp8 features"; timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.5 --feat-format mxfp8 2>gpurun_out/b5c.err | one; tail -2 gpurun_out/b5c.err
echo "== bench gather-bwd"; GLT_B200_GATHER_BWD=1 timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.5 2>/dev/null | one
echo "== sections"; timeout -k 10 200 python bench.py --sections 2>&1 | tail -1
