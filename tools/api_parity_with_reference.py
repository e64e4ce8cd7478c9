"""Compare this package's public surface with the INSTALLED reference (baseline/_ref): names exported by each
sub-package, methods of the classes both define, and keyword names of every common function / constructor / method.

  python tools/api_parity_with_reference.py        # prints the differences; 'API PARITY OK' when only the known
                                                   # internal helpers differ
"""
import importlib
import inspect
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'baseline', 'shims'))
sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))

# internals of the reference that have no public use here (steps of its all2all exchange)
KNOWN = {'distributed.DistFeature.communicate_node_feats',
         'distributed.DistFeature.communicate_node_id', 'distributed.DistFeature.remote_selecting_get_all2all',
         'distributed.DistFeature.remote_selecting_prepare'}
SUBS = ['data', 'sampler', 'loader', 'channel', 'partition', 'distributed', 'utils', 'typing']


def params(f):
  try:
    return [p for p in inspect.signature(f).parameters if p not in ('self', 'args', 'kwargs')]
  except (TypeError, ValueError):
    return None


def main():
  diffs = []
  for sname in SUBS:
    r = importlib.import_module('graphlearn_torch.' + sname)
    o = importlib.import_module('graphlearn_for_pytorch_b200.' + sname)
    for n in sorted(dir(r)):
      if n.startswith('_'):
        continue
      ro = getattr(r, n)
      mod = getattr(ro, '__module__', '') or (ro.__name__ if inspect.ismodule(ro) else '')
      if not (mod.startswith('graphlearn_torch') or mod.startswith('py_graphlearn')):
        continue
      if not hasattr(o, n):
        diffs.append(f'{sname}.{n}')
        continue
      oo = getattr(o, n)
      if inspect.isclass(ro) and inspect.isclass(oo):
        rp, op = params(ro.__init__), params(oo.__init__)
        if rp is not None and op is not None:
          diffs += [f'{sname}.{n}.__init__({p}=)' for p in rp if p not in op]
        for mn, m in inspect.getmembers(ro, predicate=inspect.isfunction):
          if mn.startswith('_'):
            continue
          om = getattr(oo, mn, None)
          if om is None:
            diffs.append(f'{sname}.{n}.{mn}')
            continue
          rp, op = params(m), params(om)
          if rp is not None and op is not None:
            diffs += [f'{sname}.{n}.{mn}({p}=)' for p in rp if p not in op]
      elif inspect.isfunction(ro) and callable(oo):
        rp, op = params(ro), params(oo)
        if rp is not None and op is not None:
          diffs += [f'{sname}.{n}({p}=)' for p in rp if p not in op]
  unknown = sorted(set(d for d in diffs if d not in KNOWN))
  for d in unknown:
    print('MISSING', d)
  print(f'{len(set(diffs)) - len(unknown)} known internal differences')
  if not unknown:
    print('API PARITY OK')
  return 1 if unknown else 0


if __name__ == '__main__':
  sys.exit(main())
