"""OGB on-disk layouts (text, .npz, heterogeneous) parse without the `ogb` package, and the examples train from them."""
import os
import subprocess
import sys

import torch

import graphlearn_for_pytorch_b200 as glt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tiny(n=60, e=400, f=8, c=5, seed=0):
  g = torch.Generator().manual_seed(seed)
  ei = torch.randint(0, n, (2, e), generator=g)
  x = torch.randn(n, f, generator=g)
  y = torch.randint(0, c, (n,), generator=g)
  perm = torch.randperm(n, generator=g)
  return ei, x, y, {'train': perm[:36], 'valid': perm[36:48], 'test': perm[48:]}


def test_text_and_binary_layouts_roundtrip(tmp_path):
  ei, x, y, split = _tiny()
  for binary, name in ((False, 'ogbn-tiny'), (True, 'ogbn-tinybin')):
    d = glt.utils.write_ogb_node_dataset(str(tmp_path), name, ei, x, y, split, split_scheme='time', binary=binary)
    assert os.path.basename(d) == name.replace('-', '_')
    for use in ('parse', 'cache'):
      o = glt.utils.load_ogb_node_dataset(str(tmp_path), name)
      assert torch.equal(o['edge_index'], ei) and torch.allclose(o['x'], x, atol=1e-6), (binary, use)
      assert o['y'].dtype == torch.int64 and torch.equal(o['y'], y) and o['num_nodes'] == x.shape[0]
      assert all(torch.equal(o['split'][k], v) for k, v in split.items())
    assert os.path.exists(os.path.join(d, 'glt_cache.pt'))
  o = glt.utils.load_ogb_node_dataset(str(tmp_path), 'ogbn-tiny', feat_dtype=torch.float16)
  assert o['x'].dtype == torch.float16


def test_unlabelled_nodes_become_minus_one(tmp_path):
  import numpy as np
  ei, x, y, split = _tiny()
  d = glt.utils.write_ogb_node_dataset(str(tmp_path), 'ogbn-p', ei, x, y, split, binary=True)
  lab = y.numpy().reshape(-1, 1).astype(np.float32)
  lab[::3] = np.nan                                     # papers100M marks unlabelled nodes with NaN
  np.savez(os.path.join(d, 'raw', 'node-label.npz'), node_label=lab)
  o = glt.utils.load_ogb_node_dataset(str(tmp_path), 'ogbn-p', use_cache=False)
  assert o['y'].dtype == torch.int64 and bool((o['y'][::3] == -1).all()) and torch.equal(o['y'][1::3], y[1::3])


def test_hetero_layout_roundtrip(tmp_path):
  g = torch.Generator().manual_seed(1)
  edges = {('a', 'to', 'b'): torch.randint(0, 12, (2, 50), generator=g),
           ('a', 'self', 'a'): torch.randint(0, 12, (2, 30), generator=g)}
  x, y = {'a': torch.randn(12, 4, generator=g)}, {'a': torch.randint(0, 3, (12,), generator=g)}
  split = {'train': {'a': torch.arange(8)}, 'test': {'a': torch.arange(8, 12)}}
  glt.utils.write_ogb_hetero_dataset(str(tmp_path), 'ogbn-h', edges, x, y, {'a': 12, 'b': 12}, split)
  for _ in range(2):
    o = glt.utils.load_ogb_hetero_dataset(str(tmp_path), 'ogbn-h')
    assert set(o['edge_index']) == set(edges) and all(torch.equal(o['edge_index'][k], v) for k, v in edges.items())
    assert o['num_nodes'] == {'a': 12, 'b': 12} and list(o['x']) == ['a'] and torch.equal(o['y']['a'], y['a'])
    assert torch.equal(o['split']['test']['a'], torch.arange(8, 12))


def _run(args, timeout=300):
  env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get('CUDA_VISIBLE_DEVICES', ''))
  out = subprocess.run([sys.executable] + args, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)
  assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
  return out.stdout


def test_examples_train_from_ogb_directories(tmp_path):
  sys.path.insert(0, os.path.join(ROOT, 'examples'))
  from common import synthetic_homo, synthetic_mag
  root = str(tmp_path)
  ei, x, y = synthetic_homo(3000, 30000, feat_dim=16, num_classes=6)
  perm = torch.randperm(3000, generator=torch.Generator().manual_seed(0))
  glt.utils.write_ogb_node_dataset(root, 'ogbn-products', ei[:, :15000], x, y,
                                   {'train': perm[:1500], 'valid': perm[1500:2000], 'test': perm[2000:]},
                                   split_scheme='sales_ranking')
  out = _run(['examples/train_sage_products.py', '--root', root, '--epochs', '1', '--batch', '256'])
  assert 'test acc' in out
  parts = str(tmp_path / 'parts')
  _run(['examples/distributed/partition_dataset.py', '--out', parts, '--root', root, '--parts', '2'])
  assert os.path.exists(os.path.join(parts, 'test_idx.pt'))
  edges, feats, labels, sizes = synthetic_mag(1500, 1000, 40, 30, feat_dim=16, num_classes=4)
  fwd = {k: v for k, v in edges.items() if not k[1].startswith('rev_')}
  p = torch.randperm(1500, generator=torch.Generator().manual_seed(0))
  glt.utils.write_ogb_hetero_dataset(root, 'ogbn-mag', fwd, {'paper': feats['paper']}, labels, sizes,
                                     {'train': {'paper': p[:900]}, 'valid': {'paper': p[900:1200]}}, split_scheme='time')
  out = _run(['examples/hetero/train_hgt_mag.py', '--root', root, '--epochs', '1', '--batch', '256'])
  assert 'Val:' in out
