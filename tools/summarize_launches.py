"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table."""
import collections
import csv
import re
import sys


def main(path, steps=1):
  with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
  agg = collections.OrderedDict()
  n = 0
  for row in csv.DictReader(lines):
    try:
      t = float(row['Metric Value'].replace(',', ''))
    except (ValueError, KeyError):
      continue
    unit = row.get('Metric Unit', 'ns')
    t = t / 1000.0 if unit in ('ns', 'nsecond') else (t * 1000.0 if unit in ('ms', 'msecond') else t)
    name = re.sub(r'^void ', '', row['Kernel Name'])
    name = re.sub(r'\(.*', '', name)[:90]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
    n += 1
  tot = sum(v[1] for v in agg.values())
  print(f'launches={n} total={tot:.1f} us  per-step={tot / steps:.1f} us ({steps} steps)')
  print(f'{"us/launch":>10} {"n/step":>7} {"us/step":>9} {"share":>6}  kernel')
  for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{v[1] / v[0]:10.1f} {v[0] / steps:7.1f} {v[1] / steps:9.1f} {100 * v[1] / tot:5.1f}%  {k}')


if __name__ == '__main__':
  main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
