"""Offline partitioning time: this library vs the unmodified reference (baseline/_ref) on the same synthetic graph
(default 1 M nodes, 20 M edges, 64 fp32 features per node, 4 partitions), written to a scratch directory.

  python benchmarks/bench_partition_cpu.py ours | reference [random | frequency]

Partitioning is the step users of the reference wait for before any distributed run (examples/distributed/
partition_ogbn_dataset.py, examples/igbh/partition.py); both arms produce the same on-disk format
(tools/partition_format_compat.py).
"""
import json
import os
import shutil
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
impl = sys.argv[1] if len(sys.argv) > 1 else 'ours'
kind = sys.argv[2] if len(sys.argv) > 2 else 'random'
N, E, F, P = int(os.environ.get('N', 1_000_000)), int(os.environ.get('E', 20_000_000)), 64, 4
g = torch.Generator().manual_seed(0)
src = torch.randint(0, N, (E,), generator=g)
dst = (src + torch.randint(1, 5000, (E,), generator=g)) % N
ei = torch.stack([src, dst])
x = torch.randn(N, F, generator=g)
if impl == 'reference':
  sys.path.insert(0, os.path.join(ROOT, 'baseline', 'shims'))
  sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))
  import graphlearn_torch as glt
else:
  import graphlearn_for_pytorch_b200 as glt
out = tempfile.mkdtemp(prefix='glt_part_')
try:
  t0 = time.time()
  if kind == 'random':
    p = glt.partition.RandomPartitioner(out, P, N, ei, node_feat=x, chunk_size=100_000)
  else:
    probs = [torch.rand(N, generator=g) for _ in range(P)]
    p = glt.partition.FrequencyPartitioner(out, P, N, ei, probs, node_feat=x, cache_ratio=0.1, chunk_size=100_000)
  p.partition()
  dt = time.time() - t0
  size = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(out) for f in fs)
  print(json.dumps({'impl': impl, 'partitioner': kind, 'nodes': N, 'edges': E, 'feat_dim': F, 'parts': P,
                    'seconds': round(dt, 2), 'M_edges_per_s': round(E / dt / 1e6, 2), 'bytes_written': size,
                    'cpu_threads': torch.get_num_threads()}))
finally:
  shutil.rmtree(out, ignore_errors=True)
