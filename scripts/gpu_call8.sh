#!/bin/bash
# GPU pass 8 (1 GPU): L2-prefetch A/B, in-kernel timeline, compute-sanitizer logs, feature micro-benchmark reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
one() { python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'l1', d['details']['layer1_autotune_ms'])"; }
echo "== fused tests"; timeout -k 10 300 python -m pytest tests/test_gpu_engine.py -m gpu -q -x > gpurun_out/pytest_eng8.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_eng8.log
echo "== bench prefetch on"; timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.5 2>/dev/null | one
echo "== bench prefetch off"; GLT_B200_L2_PREFETCH=0 timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.5 2>/dev/null | one
echo "== timeline prefetch on"; GLT_B200_FUSED_TRACE=1 timeout -k 10 300 python tools/fused_trace.py 2>/dev/null > gpurun_out/fused_trace_r2_prefetch.txt; cat gpurun_out/fused_trace_r2_prefetch.txt
echo "== timeline prefetch off"; GLT_B200_L2_PREFETCH=0 GLT_B200_FUSED_TRACE=1 timeout -k 10 300 python tools/fused_trace.py 2>/dev/null > gpurun_out/fused_trace_r2_noprefetch.txt; cat gpurun_out/fused_trace_r2_noprefetch.txt
echo "== bench_feature ref"; timeout -k 10 400 python benchmarks/bench_feature.py --impl reference 2>&1 | tail -1
echo "== sanitizer memcheck"; timeout -k 10 900 bash tools/sanitize.sh memcheck > gpurun_out/sanitizer_memcheck.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/sanitizer_memcheck.log
echo "== sanitizer racecheck"; timeout -k 10 900 bash tools/sanitize.sh racecheck > gpurun_out/sanitizer_racecheck.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/sanitizer_racecheck.log
