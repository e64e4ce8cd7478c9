#!/usr/bin/env bash
# ONE gpurun call that produces every piece of evidence profiles/ needs (each call costs ~100 s of fixed overhead,
# so batch):   gpurun --timeout 900 -- 'bash tools/gpu_round.sh [tag]'
# Writes under gpurun_out/<tag>_*; summarise locally with  python tools/summarize_round.py <tag>
set -uo pipefail
TAG=${1:-round}
OUT=gpurun_out
mkdir -p $OUT
run() { echo "== $*" >&2; timeout "${T:-200}" "$@"; }

T=300 run python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -2 $OUT/${TAG}_pytest_gpu.log
run python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
run python bench.py --steps 100 --warmup 10 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.log; tail -c 300 $OUT/${TAG}_bench.json
T=300 run python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.log
run python bench.py --sections 2> /dev/null | tail -1 > $OUT/${TAG}_sections.json
GLT_B200_FUSED_TRACE=1 run python tools/fused_trace.py --steps 30 2>&1 | grep -v Warning | tail -12 > $OUT/${TAG}_fused_trace.txt
# launch list (cold caches, serialized) and one full capture of the top kernel: never benchmark values
run ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --fused on --no-pipeline --profile-steps 4 > $OUT/${TAG}_launches.log 2>&1
run ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_sage_fused3 -c 1 \
    -o $OUT/${TAG}_prof_fused -f python bench.py --fused on --no-pipeline --profile-steps 1 > $OUT/${TAG}_ncu.log 2>&1
run python benchmarks/bench_sampler.py 2>/dev/null | tail -1 > $OUT/${TAG}_bench_sampler.json
run python benchmarks/bench_feature.py 2>/dev/null | tail -1 > $OUT/${TAG}_bench_feature.json
if [[ "${GLT_B200_EXPERIMENTAL:-0}" == "1" ]]; then
  GLT_B200_EXPERIMENTAL=1 run python -m pytest tests/test_gpu_engine.py -q -k "transposed or gather_backward or engine_and_trainer" > $OUT/${TAG}_experimental.log 2>&1
  tail -3 $OUT/${TAG}_experimental.log
  GLT_B200_GATHER_BWD=1 run python bench.py --steps 100 --warmup 10 > $OUT/${TAG}_bench_gather_bwd.json 2> /dev/null
  GLT_B200_AGG_BATCH=1 run python -m pytest tests/test_gpu_engine.py -q -k "forward_backward" > $OUT/${TAG}_agg_batch_test.log 2>&1; tail -1 $OUT/${TAG}_agg_batch_test.log
  GLT_B200_AGG_BATCH=1 run python bench.py --steps 100 --warmup 10 > $OUT/${TAG}_bench_agg_batch.json 2> /dev/null
  tail -c 200 $OUT/${TAG}_bench_gather_bwd.json
fi
echo "done: $(ls $OUT | grep -c "^${TAG}_") files"
