"""Helpers shared by the example scripts: synthetic datasets with the shapes of the public
benchmarks (no network access here; swap in real tensors where marked)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphlearn_for_pytorch_b200 as glt  # noqa: E402
from graphlearn_for_pytorch_b200.utils.synthetic import rmat_edges  # noqa: E402


def synthetic_homo(num_nodes=100_000, num_edges=2_000_000, feat_dim=100, num_classes=47, seed=0,
                   learnable=True):
  """(edge_index [2,E] undirected, features fp32, labels).  With learnable=True the labels are a
  linear function of the features so that accuracy is meaningful."""
  ei = rmat_edges(num_nodes, num_edges // 2, seed=seed)
  ei = torch.cat([ei, ei.flip(0)], 1)
  g = torch.Generator().manual_seed(seed + 1)
  x = torch.randn(num_nodes, feat_dim, generator=g)
  if learnable:
    w = torch.randn(feat_dim, num_classes, generator=g)
    y = (x @ w).argmax(1)
  else:
    y = torch.randint(0, num_classes, (num_nodes,), generator=g)
  return ei, x, y


def synthetic_igbh(num_papers=20_000, num_authors=10_000, num_insts=500, num_fos=200, feat_dim=128,
                   num_classes=19, seed=0):
  """IGBH-shaped hetero graph (reference examples/igbh/dataset.py:153-166): paper/author/institute/fos
  with cites, written_by, affiliated_to, topic (+ reverse relations)."""
  g = torch.Generator().manual_seed(seed)

  def rnd(n_src, n_dst, n_e):
    return torch.stack([torch.randint(0, n_src, (n_e,), generator=g), torch.randint(0, n_dst, (n_e,), generator=g)])
  cites = rnd(num_papers, num_papers, num_papers * 8)
  cites = torch.cat([cites, cites.flip(0)], 1)
  written = rnd(num_papers, num_authors, num_papers * 3)
  affil = rnd(num_authors, num_insts, num_authors * 2)
  topic = rnd(num_papers, num_fos, num_papers * 2)
  edges = {
    ('paper', 'cites', 'paper'): cites,
    ('paper', 'written_by', 'author'): written,
    ('author', 'rev_written_by', 'paper'): written.flip(0),
    ('author', 'affiliated_to', 'institute'): affil,
    ('institute', 'rev_affiliated_to', 'author'): affil.flip(0),
    ('paper', 'topic', 'fos'): topic,
    ('fos', 'rev_topic', 'paper'): topic.flip(0),
  }
  sizes = {'paper': num_papers, 'author': num_authors, 'institute': num_insts, 'fos': num_fos}
  feats = {t: torch.randn(n, feat_dim, generator=g) for t, n in sizes.items()}
  w = torch.randn(feat_dim, num_classes, generator=g)
  labels = {'paper': (feats['paper'] @ w).argmax(1)}
  return edges, feats, labels, sizes


def synthetic_mag(num_papers=30_000, num_authors=20_000, num_insts=500, num_fields=300, feat_dim=128,
                  num_classes=20, seed=0):
  """OGB-MAG-shaped heterogeneous graph (paper / author / institution / field_of_study with cites, writes,
  affiliated_with, has_topic, made undirected like T.ToUndirected(merge=True) does in the reference example
  examples/hetero/train_hgt_mag.py:86-91).  Every node type gets dense features (the reference fills the
  feature-less types with metapath2vec embeddings)."""
  g = torch.Generator().manual_seed(seed)

  def rnd(n_src, n_dst, n_e):
    return torch.stack([torch.randint(0, n_src, (n_e,), generator=g), torch.randint(0, n_dst, (n_e,), generator=g)])
  cites = rnd(num_papers, num_papers, num_papers * 6)
  writes = rnd(num_authors, num_papers, num_papers * 3)
  affil = rnd(num_authors, num_insts, num_authors)
  topic = rnd(num_papers, num_fields, num_papers * 4)
  edges = {
    ('paper', 'cites', 'paper'): torch.cat([cites, cites.flip(0)], 1),
    ('author', 'writes', 'paper'): writes,
    ('paper', 'rev_writes', 'author'): writes.flip(0),
    ('author', 'affiliated_with', 'institution'): affil,
    ('institution', 'rev_affiliated_with', 'author'): affil.flip(0),
    ('paper', 'has_topic', 'field_of_study'): topic,
    ('field_of_study', 'rev_has_topic', 'paper'): topic.flip(0),
  }
  sizes = {'paper': num_papers, 'author': num_authors, 'institution': num_insts, 'field_of_study': num_fields}
  feats = {t: torch.randn(n, feat_dim, generator=g) for t, n in sizes.items()}
  w = torch.randn(feat_dim, num_classes, generator=g)
  labels = {'paper': (feats['paper'] @ w).argmax(1)}
  return edges, feats, labels, sizes


def add_dataset_args(parser, nodes=100_000, edges=2_000_000):
  """--root / --dataset select an OGB directory on disk (parsed by glt.utils.load_ogb_node_dataset, no `ogb` package
  needed); without --root a synthetic graph of --nodes / --edges is generated."""
  parser.add_argument('--root', default=None, help='directory that contains e.g. ogbn_products/ (OGB raw layout)')
  parser.add_argument('--dataset', default='ogbn-products')
  parser.add_argument('--nodes', type=int, default=nodes)
  parser.add_argument('--edges', type=int, default=edges)


def load_homo(args, feat_dim=100, num_classes=47, undirected=True):
  """-> (edge_index, x fp32, y int64, {'train' | 'valid' | 'test': indices}, num_nodes)."""
  if getattr(args, 'root', None):
    d = glt.utils.load_ogb_node_dataset(args.root, args.dataset)
    ei, n = d['edge_index'], d['num_nodes']
    if undirected:                                      # what T.ToUndirected / to_symmetric do in the reference scripts
      ei = torch.cat([ei, ei.flip(0)], 1)
    y = d['y'].to(torch.int64)
    split = dict(d['split'])
    if 'train' not in split:
      perm = torch.randperm(n, generator=torch.Generator().manual_seed(0))
      split = {'train': perm[: n // 10], 'valid': perm[n // 10: n // 10 + n // 50], 'test': perm[-(n // 50):]}
    return ei, d['x'], y, split, n
  ei, x, y = synthetic_homo(args.nodes, args.edges, feat_dim=feat_dim, num_classes=num_classes)
  n = args.nodes
  perm = torch.randperm(n, generator=torch.Generator().manual_seed(0))
  n_tr, n_ev = n // 10, max(n // 50, 1)
  return ei, x, y, {'train': perm[:n_tr], 'valid': perm[n_tr:n_tr + n_ev], 'test': perm[n_tr + n_ev:n_tr + 2 * n_ev]}, n
