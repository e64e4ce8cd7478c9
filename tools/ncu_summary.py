"""Extract the judge-relevant metrics from an ncu report into a text summary.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
  'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
  'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
  'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
  'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
  'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
  'sm__inst_executed_pipe_tensor', 'smsp__cycles_active.avg',
  'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
  'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
  'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
  'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
  'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
  'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
]


def main(path):
  raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  hdr, units = rows[0], rows[1]
  peaks = {}
  try:
    peaks = json.load(open('MEASURED_PEAKS.json'))
  except Exception:
    pass
  for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    print(f'== {name}')
    vals = {}
    for i, h in enumerate(hdr):
      if any(h == k or h.startswith(k) for k in KEYS) and r[i] not in ('',):
        print(f'   {h:85s} {r[i]:>16s} {units[i]}')
        vals[h] = (r[i], units[i])
    try:
      t = float(vals['gpu__time_duration.sum'][0].replace(',', ''))
      tu = vals['gpu__time_duration.sum'][1]
      t_s = t * {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1}.get(tu, 1e-9)
      def to_bytes(k):
        v, u = vals[k]
        return float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
      traffic = to_bytes('dram__bytes_read.sum') + to_bytes('dram__bytes_write.sum')
      bw = traffic / t_s / 1e9
      line = f'   -> DRAM traffic {traffic / 1e6:.1f} MB in {t_s * 1e6:.1f} us = {bw:.0f} GB/s'
      if peaks.get('hbm_gbs'):
        line += f' = {100 * bw / peaks["hbm_gbs"]:.1f}% of measured copy peak ({peaks["hbm_gbs"]} GB/s)'
      print(line)
    except Exception:
      pass


if __name__ == '__main__':
  main(sys.argv[1])
