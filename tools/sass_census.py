"""SASS instruction census per kernel of the built extension (proves tcgen05 / TMA / mbarrier use).

  python tools/sass_census.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'graphlearn_for_pytorch_b200', '_ext',
                  'glt_b200_C.so')
KEYS = ['UTCHMMA', 'LDTM', 'UBLKCP', 'UTCBAR', 'UTCATOMSWS', 'SYNCS', 'LDG.E', 'STG.E', 'LDS', 'STS', 'ATOMG', 'RED', 'STL',
        'LDL', 'MULTIMEM']


def main():
  out = subprocess.run(['cuobjdump', '-sass', SO], capture_output=True, text=True).stdout
  print(f'SASS instruction census per kernel of {os.path.relpath(SO)} (cuobjdump -sass, sm_100a)')
  print('UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA bulk copy), UTCBAR = tcgen05.commit, '
        'UTCATOMSWS = tcgen05.alloc/dealloc, SYNCS = mbarrier ops, STL/LDL = register spills\n')
  name, counts, total = None, collections.Counter(), 0

  def flush():
    if name:
      print(name)
      print('    total=%d %s' % (total, ' '.join(f'{k}={v}' for k, v in counts.items() if v)))

  for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
      flush()
      name, counts, total = m.group(1), collections.Counter(), 0
      continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
    if m:
      total += 1
      op = m.group(1)
      for k in KEYS:
        if op.startswith(k):
          counts[k] += 1
          break
  flush()


if __name__ == '__main__':
  main()
