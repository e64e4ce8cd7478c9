"""Run the REFERENCE's own python unit tests against this package (drop-in check).

`import graphlearn_torch` is aliased to graphlearn_for_pytorch_b200 in a scratch directory, the reference's
test/python/*.py are copied next to it and run with pytest.  Most of them build CUDA graphs in setUp, so the useful
run is on a GPU box; on a CPU-only machine test_partition / test_graph / the CPU cases of the sampler tests run.

  python tools/run_reference_tests.py [/root/reference] [--patch-cuda-to-cpu] [-k expr]

--patch-cuda-to-cpu rewrites `torch.device('cuda:0')` / `torch.device('cuda', 0)` in the copied tests to the CPU
device, which lets the sampler tests' logic (expected node lists, edge indices, edge ids) run on a CPU-only machine.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALIAS = '''import sys
import graphlearn_for_pytorch_b200 as _impl
sys.modules[__name__] = _impl
for _m in list(sys.modules):
  if _m.startswith('graphlearn_for_pytorch_b200.'):
    sys.modules['graphlearn_torch.' + _m[len('graphlearn_for_pytorch_b200.'):]] = sys.modules[_m]
'''
# `parameterized` is not installed in the image: the two decorators the reference tests use, in 20 lines
PARAMETERIZED = '''import functools


class parameterized(object):
  @staticmethod
  def expand(cases):
    def deco(fn):
      import sys
      frame = sys._getframe(1)
      for i, case in enumerate(cases):
        args = case if isinstance(case, (list, tuple)) else (case,)

        def make(a):
          @functools.wraps(fn)
          def test(self):
            return fn(self, *a)
          return test
        t = make(tuple(args))
        t.__name__ = f'{fn.__name__}_{i}'
        frame.f_locals[t.__name__] = t
      return None
    return deco
'''
SKIP = {'test_vineyard.py',               # needs a vineyard server
        'test_sample_prob.py'}            # downloads OGB-MAG through torch_geometric.datasets


def main():
  args = sys.argv[1:]
  ref = args.pop(0) if args and not args[0].startswith('-') else '/root/reference'
  patch = '--patch-cuda-to-cpu' in args
  if patch:
    args.remove('--patch-cuda-to-cpu')
  only = None
  for a in list(args):
    if a.startswith('--only='):       # --only=test_subgraph.py,test_feature.py
      only = set(a[len('--only='):].split(','))
      args.remove(a)
  show = '--show' in args             # print the whole pytest output of every file
  if show:
    args.remove('--show')
  src = os.path.join(ref, 'test', 'python')
  if not os.path.isdir(src):
    print('no reference tests at', src)
    return 0
  per_file_timeout = int(os.environ.get('GLT_REFTEST_TIMEOUT', '240'))
  work = tempfile.mkdtemp(prefix='glt_b200_reftests_')
  os.makedirs(os.path.join(work, 'alias', 'graphlearn_torch'))
  with open(os.path.join(work, 'alias', 'graphlearn_torch', '__init__.py'), 'w') as f:
    f.write(ALIAS)
  with open(os.path.join(work, 'alias', 'parameterized.py'), 'w') as f:
    f.write(PARAMETERIZED)
  # the tests build / check torch_geometric Data objects: map them to this package's attribute-compatible containers
  os.makedirs(os.path.join(work, 'alias', 'torch_geometric', 'data'))
  with open(os.path.join(work, 'alias', 'torch_geometric', '__init__.py'), 'w') as f:
    f.write('from . import data, utils\n')
  # test_pyg_remote_backend uses two index helpers of torch_geometric.utils.sparse
  os.makedirs(os.path.join(work, 'alias', 'torch_geometric', 'utils'))
  with open(os.path.join(work, 'alias', 'torch_geometric', 'utils', '__init__.py'), 'w') as f:
    f.write('from . import sparse\n')
  with open(os.path.join(work, 'alias', 'torch_geometric', 'utils', 'sparse.py'), 'w') as f:
    f.write('import torch\n\n\n'
            'def index2ptr(index, size=None):\n'
            '  size = int(index.max()) + 1 if size is None else size\n'
            '  return torch.cat([index.new_zeros(1), torch.bincount(index, minlength=size).cumsum(0)])\n\n\n'
            'def ptr2index(ptr):\n'
            '  return torch.repeat_interleave(torch.arange(ptr.numel() - 1, device=ptr.device), ptr[1:] - ptr[:-1])\n')
  with open(os.path.join(work, 'alias', 'torch_geometric', 'data', '__init__.py'), 'w') as f:
    f.write('from graphlearn_for_pytorch_b200.loader.data import Data, HeteroData  # noqa: F401\n')
  tests = os.path.join(work, 'tests')
  shutil.copytree(src, tests)
  files = sorted(f for f in os.listdir(tests) if f.startswith('test_') and f.endswith('.py') and f not in SKIP
                 and (only is None or f in only))
  if patch:
    for f in files:
      path = os.path.join(tests, f)
      src_txt = open(path).read()
      # every spelling of a CUDA device (cuda:0 / ('cuda', i % n) / bare 'cuda'), and a device count of at least one
      src_txt = re.sub(r"torch\.device\(\s*'cuda(:\d+)?'\s*(,[^()]*(\([^()]*\))?[^()]*)?\)", "torch.device('cpu')", src_txt)
      src_txt = re.sub(r"'cuda(:\d+)?'", "'cpu'", src_txt)                      # device='cuda:1'
      src_txt = src_txt.replace('torch.cuda.device_count()', 'max(torch.cuda.device_count(), 1)')
      src_txt = re.sub(r"device\s*=\s*0\b", "device='cpu'", src_txt)           # torch.tensor(..., device=0)
      src_txt = re.sub(r"\.to\(0\)", ".to('cpu')", src_txt).replace('.cuda()', '.cpu()')
      open(path, 'w').write(src_txt)
  env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(work, 'alias'), ROOT,
                                                     os.path.join(ROOT, 'baseline', 'shims')]))
  rc = 0
  for f in files:
    try:
      out = subprocess.run([sys.executable, '-m', 'pytest', f, '-q', '-p', 'no:cacheprovider'] + args, cwd=tests,
                           env=env, capture_output=True, text=True, timeout=per_file_timeout)
    except subprocess.TimeoutExpired:
      print(f'{f:40s} TIMEOUT after {per_file_timeout} s (tests that spawn CUDA sampling workers hang without a GPU)')
      rc |= 1
      continue
    tail = [ln for ln in out.stdout.strip().splitlines() if 'passed' in ln or 'failed' in ln or 'error' in ln]
    print(f'{f:40s} {tail[-1] if tail else out.stdout[-200:]}', flush=True)
    if show:
      print(out.stdout[-6000:], out.stderr[-3000:], flush=True)
    rc |= out.returncode
  return rc


if __name__ == '__main__':
  sys.exit(main())
