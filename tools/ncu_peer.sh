#!/bin/bash
# NVLink counters of the in-kernel peer loads (2 GPUs): rank 1 passive, rank 0 under ncu.  Never wraps a whole
# multi-rank launcher in ncu -- the profiled process is a single rank whose kernels need nothing from the peer but
# its (static) memory.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29733 WORLD_SIZE=2 NCCL_DEBUG_FILE=/dev/stderr
ncu --query-metrics 2>/dev/null | grep -iE "^nvl(rx|tx)__bytes" | head -8 > gpurun_out/nvlink_metric_names.txt
RANK=1 LOCAL_RANK=1 python tools/peer_passive.py > gpurun_out/peer_passive_rank1.log 2>&1 &
P1=$!
RANK=0 LOCAL_RANK=0 timeout -k 10 600 ncu --profile-from-start off --clock-control none \
  --metrics gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  -k 'regex:k_sample_hop|k_sage_fused3|k_gather_vec' -c 15 --csv --log-file gpurun_out/nvlink_peer_kernels.csv \
  python tools/peer_passive.py > gpurun_out/peer_passive_rank0.log 2>&1
echo "rank0 rc=$?"
wait $P1
tail -2 gpurun_out/peer_passive_rank0.log
python - <<'PY'
import csv
rows = [l for l in open('gpurun_out/nvlink_peer_kernels.csv') if not l.startswith('==')]
agg = {}
for r in csv.DictReader(rows):
  k = (r['ID'], r['Kernel Name'].split('(')[0][-40:])
  agg.setdefault(k, {})[r['Metric Name']] = (r['Metric Value'], r['Metric Unit'])
for (i, name), m in agg.items():
  print(i, name, {k: v for k, v in m.items()})
PY
