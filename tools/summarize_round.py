"""Turn the raw outputs of tools/gpu_round.sh (gpurun_out/<tag>_*) into the committed summaries under profiles/.

  python tools/summarize_round.py <tag> [--name r2]
"""
import argparse
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('tag')
  ap.add_argument('--name', default=None, help='suffix of the files written to profiles/ (default: the tag)')
  a = ap.parse_args()
  name = a.name or a.tag
  src = lambda f: os.path.join(ROOT, 'gpurun_out', f'{a.tag}_{f}')          # noqa: E731
  dst = lambda f: os.path.join(ROOT, 'profiles', f)                          # noqa: E731
  done = []
  for f, out in (('bench.json', f'bench_{name}_1gpu.json'), ('bench_ref.json', f'bench_{name}_reference_1gpu.json'),
                 ('sections.json', f'sections_{name}.json'), ('fused_trace.txt', f'fused_trace_{name}.txt'),
                 ('bench_sampler.json', f'bench_sampler_{name}.json'), ('bench_feature.json', f'bench_feature_{name}.json'),
                 ('bench_gather_bwd.json', f'bench_{name}_gather_bwd.json')):
    if os.path.exists(src(f)) and os.path.getsize(src(f)) > 0:
      shutil.copy(src(f), dst(out)); done.append(out)
  if os.path.exists(src('launches.csv')):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'summarize_launches.py'), src('launches.csv'), '4'],
                         capture_output=True, text=True).stdout
    open(dst(f'launches_{name}_step_breakdown.txt'), 'w').write(txt); done.append(f'launches_{name}_step_breakdown.txt')
  if os.path.exists(src('prof_fused.ncu-rep')):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_summary.py'), src('prof_fused.ncu-rep')],
                         capture_output=True, text=True).stdout
    open(dst(f'ncu_sage_fused_{name}.txt'), 'w').write(txt); done.append(f'ncu_sage_fused_{name}.txt')
  for f in ('bench.json', 'bench_ref.json', 'bench_gather_bwd.json'):
    if os.path.exists(src(f)) and os.path.getsize(src(f)) > 0:
      try:
        d = json.loads(open(src(f)).read().strip().splitlines()[-1])
        print(f, d.get('impl', 'ours'), d.get('ms_per_step'), 'ms/step', round(d.get('value', 0)), d.get('unit'))
      except Exception as e:  # noqa: BLE001
        print(f, 'unreadable:', e)
  print('wrote', done)


if __name__ == '__main__':
  main()
