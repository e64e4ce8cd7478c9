"""Readers for datasets in the on-disk layout of the Open Graph Benchmark downloader -- without the `ogb` package.

The reference's examples obtain their data through `ogb.nodeproppred.PygNodePropPredDataset`
(examples/train_sage_ogbn_products.py:95-101, examples/distributed/partition_ogbn_dataset.py:33-60,
examples/multi_gpu/train_sage_ogbn_papers100m.py:99-108).  A user who already has `<root>/ogbn_products/` on disk can
point every example of this repository at it (`--root <root> --dataset ogbn-products`); nothing but pandas / numpy is
needed to parse it.

Layouts understood (directory name = dataset name with '-' replaced by '_'):

  homogeneous, text      raw/edge.csv.gz (src,dst per line)  raw/num-node-list.csv.gz  raw/node-feat.csv.gz
  (products, arxiv)      raw/node-label.csv.gz  [raw/edge-feat.csv.gz]  split/<scheme>/{train,valid,test}.csv.gz
  homogeneous, binary    raw/data.npz {edge_index, node_feat, num_nodes_list, [edge_feat]}  raw/node-label.npz
  (papers100M)           {node_label}  split/<scheme>/{train,valid,test}.csv.gz
  heterogeneous, text    raw/triplet-type-list.csv.gz  raw/num-node-dict.csv.gz  raw/relations/<s>___<r>___<d>/edge.csv.gz
  (mag)                  raw/node-feat/<type>/node-feat.csv.gz  raw/node-label/<type>/node-label.csv.gz
                         split/<scheme>/<type>/{train,valid,test}.csv.gz

The parsed tensors are cached next to the raw files (`glt_cache.pt`, plain tensors, `weights_only`-loadable), so the
text parse happens once.  `write_ogb_node_dataset` / `write_ogb_hetero_dataset` produce the same layouts from tensors
(used by the tests; also handy to hand a synthetic graph to tooling that expects OGB directories).
"""
import gzip
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

__all__ = ['load_ogb_node_dataset', 'load_ogb_hetero_dataset', 'write_ogb_node_dataset', 'write_ogb_hetero_dataset',
           'ogb_dir_name']

_CACHE = 'glt_cache.pt'
_SPLITS = ('train', 'valid', 'test')


def ogb_dir_name(name: str) -> str:
  return name.replace('-', '_')


def _read_csv(path: str, dtype) -> np.ndarray:
  """One comma-separated record per line, no header (pandas' C parser; falls back to numpy when pandas is absent)."""
  try:
    import pandas as pd
    return pd.read_csv(path, compression='gzip' if path.endswith('.gz') else None, header=None).values.astype(dtype)
  except ImportError:  # pragma: no cover
    with (gzip.open(path, 'rt') if path.endswith('.gz') else open(path)) as f:
      return np.loadtxt(f, delimiter=',', dtype=dtype, ndmin=2)


def _write_csv(path: str, arr: np.ndarray, fmt: str):
  os.makedirs(os.path.dirname(path), exist_ok=True)
  arr = np.asarray(arr)
  if arr.ndim == 1:
    arr = arr[:, None]
  with gzip.open(path, 'wt') as f:
    np.savetxt(f, arr, delimiter=',', fmt=fmt)


def _find(dirname: str, stem: str) -> Optional[str]:
  for ext in ('.csv.gz', '.csv'):
    p = os.path.join(dirname, stem + ext)
    if os.path.exists(p):
      return p
  return None


def _split_scheme(ds_dir: str, scheme: Optional[str]) -> Optional[str]:
  sdir = os.path.join(ds_dir, 'split')
  if scheme is not None:
    return os.path.join(sdir, scheme)
  if not os.path.isdir(sdir):
    return None
  found = sorted(d for d in os.listdir(sdir) if os.path.isdir(os.path.join(sdir, d)))
  return os.path.join(sdir, found[0]) if found else None


def _read_split(split_dir: Optional[str]) -> Dict[str, torch.Tensor]:
  out = {}
  if split_dir is None:
    return out
  for s in _SPLITS:
    p = _find(split_dir, s)
    if p is not None:
      out[s] = torch.from_numpy(_read_csv(p, np.int64).reshape(-1))
  return out


def _labels(arr: np.ndarray) -> torch.Tensor:
  """[N, 1] label columns become [N]; NaN (unlabelled nodes of papers100M) becomes -1; integral values -> int64."""
  a = np.asarray(arr)
  if a.ndim == 2 and a.shape[1] == 1:
    a = a[:, 0]
  if np.issubdtype(a.dtype, np.floating):
    a = np.where(np.isnan(a), -1, a)
    if np.all(a == np.round(a)):
      a = a.astype(np.int64)
  return torch.from_numpy(np.ascontiguousarray(a))


def load_ogb_node_dataset(root: str, name: str, split_scheme: Optional[str] = None, use_cache: bool = True,
                          feat_dtype: torch.dtype = torch.float32) -> Dict[str, object]:
  """Homogeneous node-property dataset -> dict with
       edge_index [2, E] int64, x [N, F] | None, y [N] | None, edge_attr | None, num_nodes,
       split {'train' | 'valid' | 'test': int64 indices}."""
  ds_dir = os.path.join(root, ogb_dir_name(name))
  raw = os.path.join(ds_dir, 'raw')
  if not os.path.isdir(raw):
    raise FileNotFoundError(f'{raw}: no OGB raw directory (expected the layout written by the OGB downloader)')
  cache = os.path.join(ds_dir, _CACHE)
  if use_cache and os.path.exists(cache):
    out = torch.load(cache, weights_only=True)
    if out['x'] is not None:
      out['x'] = out['x'].to(feat_dtype)
    return out
  npz = os.path.join(raw, 'data.npz')
  if os.path.exists(npz):                                    # binary layout (ogbn-papers100M)
    d = np.load(npz)
    ei = torch.from_numpy(np.ascontiguousarray(d['edge_index']).astype(np.int64))
    x = torch.from_numpy(np.ascontiguousarray(d['node_feat'])) if 'node_feat' in d.files else None
    ea = torch.from_numpy(np.ascontiguousarray(d['edge_feat'])) if 'edge_feat' in d.files else None
    n = int(np.asarray(d['num_nodes_list']).reshape(-1)[0])
    lab = os.path.join(raw, 'node-label.npz')
    y = _labels(np.load(lab)['node_label']) if os.path.exists(lab) else None
  else:
    ep = _find(raw, 'edge')
    if ep is None:
      raise FileNotFoundError(f'{raw}: neither data.npz nor edge.csv[.gz]')
    ei = torch.from_numpy(np.ascontiguousarray(_read_csv(ep, np.int64).T))
    n = int(_read_csv(_find(raw, 'num-node-list'), np.int64).reshape(-1)[0])
    fp, lp, efp = _find(raw, 'node-feat'), _find(raw, 'node-label'), _find(raw, 'edge-feat')
    x = torch.from_numpy(_read_csv(fp, np.float32)) if fp else None
    y = _labels(_read_csv(lp, np.float64)) if lp else None
    ea = torch.from_numpy(_read_csv(efp, np.float32)) if efp else None
  out = {'edge_index': ei, 'x': x, 'y': y, 'edge_attr': ea, 'num_nodes': n,
         'split': _read_split(_split_scheme(ds_dir, split_scheme))}
  if use_cache:
    tmp = cache + f'.tmp{os.getpid()}'
    torch.save(out, tmp)
    os.replace(tmp, cache)
  if x is not None:
    out['x'] = x.to(feat_dtype)
  return out


def load_ogb_hetero_dataset(root: str, name: str, split_scheme: Optional[str] = None, use_cache: bool = True
                            ) -> Dict[str, object]:
  """Heterogeneous node-property dataset (ogbn-mag layout) -> dict with
       edge_index {(src, rel, dst): [2, E]}, x {type: [N, F]}, y {type: [N]}, num_nodes {type: N},
       split {'train' | 'valid' | 'test': {type: indices}}."""
  ds_dir = os.path.join(root, ogb_dir_name(name))
  raw = os.path.join(ds_dir, 'raw')
  cache = os.path.join(ds_dir, _CACHE)
  if use_cache and os.path.exists(cache):
    flat = torch.load(cache, weights_only=True)
    flat['edge_index'] = {tuple(k.split('___')): v for k, v in flat['edge_index'].items()}
    return flat
  import pandas as pd
  trip = pd.read_csv(_find(raw, 'triplet-type-list'), header=None).values.tolist()
  nd = pd.read_csv(_find(raw, 'num-node-dict'))              # header row = node types, one row of counts
  num_nodes = {t: int(nd[t][0]) for t in nd.columns}
  edge_index = {}
  for s, r, d in trip:
    p = _find(os.path.join(raw, 'relations', f'{s}___{r}___{d}'), 'edge')
    edge_index[(s, r, d)] = torch.from_numpy(np.ascontiguousarray(_read_csv(p, np.int64).T))
  x, y = {}, {}
  for t in num_nodes:
    fp = _find(os.path.join(raw, 'node-feat', t), 'node-feat')
    if fp:
      x[t] = torch.from_numpy(_read_csv(fp, np.float32))
    lp = _find(os.path.join(raw, 'node-label', t), 'node-label')
    if lp:
      y[t] = _labels(_read_csv(lp, np.float64))
  split = {}
  sdir = _split_scheme(ds_dir, split_scheme)
  if sdir is not None:
    for t in num_nodes:
      for s, idx in _read_split(os.path.join(sdir, t)).items():
        split.setdefault(s, {})[t] = idx
  out = {'edge_index': edge_index, 'x': x, 'y': y, 'num_nodes': num_nodes, 'split': split}
  if use_cache:
    flat = dict(out)
    flat['edge_index'] = {'___'.join(k): v for k, v in edge_index.items()}   # string keys: weights_only-loadable
    tmp = cache + f'.tmp{os.getpid()}'
    torch.save(flat, tmp)
    os.replace(tmp, cache)
  return out


def write_ogb_node_dataset(root: str, name: str, edge_index: torch.Tensor, x: Optional[torch.Tensor],
                           y: Optional[torch.Tensor], split: Dict[str, torch.Tensor], split_scheme: str = 'random',
                           binary: bool = False, num_nodes: Optional[int] = None) -> str:
  """Write tensors in the OGB raw layout (text, or the .npz layout of papers100M with binary=True)."""
  ds_dir = os.path.join(root, ogb_dir_name(name))
  raw = os.path.join(ds_dir, 'raw')
  os.makedirs(raw, exist_ok=True)
  n = int(num_nodes if num_nodes is not None else (x.shape[0] if x is not None else int(edge_index.max()) + 1))
  if binary:
    arrs = {'edge_index': edge_index.numpy(), 'num_nodes_list': np.array([n]),
            'num_edges_list': np.array([edge_index.shape[1]])}
    if x is not None:
      arrs['node_feat'] = x.numpy()
    np.savez(os.path.join(raw, 'data.npz'), **arrs)
    if y is not None:
      np.savez(os.path.join(raw, 'node-label.npz'), node_label=y.numpy().reshape(-1, 1).astype(np.float32))
  else:
    _write_csv(os.path.join(raw, 'edge.csv.gz'), edge_index.t().numpy(), '%d')
    _write_csv(os.path.join(raw, 'num-node-list.csv.gz'), np.array([n]), '%d')
    _write_csv(os.path.join(raw, 'num-edge-list.csv.gz'), np.array([edge_index.shape[1]]), '%d')
    if x is not None:
      _write_csv(os.path.join(raw, 'node-feat.csv.gz'), x.numpy(), '%.7g')
    if y is not None:
      _write_csv(os.path.join(raw, 'node-label.csv.gz'), y.numpy(), '%d')
  for s, idx in split.items():
    _write_csv(os.path.join(ds_dir, 'split', split_scheme, f'{s}.csv.gz'), idx.numpy(), '%d')
  return ds_dir


def write_ogb_hetero_dataset(root: str, name: str, edge_index: Dict[Tuple[str, str, str], torch.Tensor],
                             x: Dict[str, torch.Tensor], y: Dict[str, torch.Tensor], num_nodes: Dict[str, int],
                             split: Dict[str, Dict[str, torch.Tensor]], split_scheme: str = 'random') -> str:
  ds_dir = os.path.join(root, ogb_dir_name(name))
  raw = os.path.join(ds_dir, 'raw')
  os.makedirs(raw, exist_ok=True)
  with gzip.open(os.path.join(raw, 'triplet-type-list.csv.gz'), 'wt') as f:
    f.writelines(','.join(k) + '\n' for k in edge_index)
  types = list(num_nodes)
  with gzip.open(os.path.join(raw, 'num-node-dict.csv.gz'), 'wt') as f:
    f.write(','.join(types) + '\n' + ','.join(str(int(num_nodes[t])) for t in types) + '\n')
  for (s, r, d), ei in edge_index.items():
    rel = os.path.join(raw, 'relations', f'{s}___{r}___{d}')
    _write_csv(os.path.join(rel, 'edge.csv.gz'), ei.t().numpy(), '%d')
    _write_csv(os.path.join(rel, 'num-edge-list.csv.gz'), np.array([ei.shape[1]]), '%d')
  for t, v in x.items():
    _write_csv(os.path.join(raw, 'node-feat', t, 'node-feat.csv.gz'), v.numpy(), '%.7g')
  for t, v in y.items():
    _write_csv(os.path.join(raw, 'node-label', t, 'node-label.csv.gz'), v.numpy(), '%d')
  for s, per_type in split.items():
    for t, idx in per_type.items():
      _write_csv(os.path.join(ds_dir, 'split', split_scheme, t, f'{s}.csv.gz'), idx.numpy(), '%d')
  return ds_dir
