#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== kernel times gather-bwd"; GLT_B200_GATHER_BWD=1 timeout -k 10 300 python bench.py --kernel-times 2>&1 | grep -v Warning | tee gpurun_out/kernel_times_insitu_gatherbwd.txt
echo "== in-degree stats"; timeout -k 10 300 python tools/indegree_stats.py 2>&1 | grep -v Warning | tail -12
