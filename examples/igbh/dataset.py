"""IGBH heterogeneous dataset on disk (the MLPerf-GNN reference workload; counterpart of the reference's
examples/igbh/dataset.py:137-285).

On-disk layout (same as the IGB release / the reference):
  <path>/<size>/processed/<src>__<rel>__<dst>/edge_index.npy          int64 [E, 2]
  <path>/<size>/processed/<ntype>/node_feat.npy                       float32 [N, F]
  <path>/<size>/processed/paper/node_label_19.npy | node_label_2K.npy int64 [N_paper]
optional products of the preprocessing scripts next to them:
  compress_graph.py -> <layout>/<etype>/{compressed_0,compressed_1}.pt (+ edge order) and node_feat_fp16.pt
  split_seeds.py    -> train_idx.pt / val_idx.pt

There is no network here, so `make_synthetic_igbh()` writes a graph with IGBH's schema (4 node types,
4 base relations, reverse relations added at load time, features correlated with the labels) in
exactly this layout; every other script consumes the files, not the generator.
"""
import os
import os.path as osp
from typing import Dict, Tuple

import numpy as np
import torch

NTYPES = ['paper', 'author', 'institute', 'fos']
BASE_ETYPES = [('paper', 'cites', 'paper'), ('paper', 'written_by', 'author'),
               ('author', 'affiliated_to', 'institute'), ('paper', 'topic', 'fos')]
# IGBH-large / -full also carry the publication venues (reference dataset.py:186-210)
VENUE_NTYPES = ['conference', 'journal']
VENUE_ETYPES = [('paper', 'published', 'journal'), ('paper', 'venue', 'conference')]


def ntypes_on_disk(base_path: str):
  return NTYPES + [nt for nt in VENUE_NTYPES if osp.exists(osp.join(base_path, nt, 'node_feat.npy'))]


def etypes_on_disk(base_path: str):
  return BASE_ETYPES + [et for et in VENUE_ETYPES if osp.isdir(osp.join(base_path, etype_dir(et)))]


def etype_dir(et: Tuple[str, str, str]) -> str:
  return '__'.join(et)


def make_synthetic_igbh(path: str, size: str = 'tiny', papers: int = 4000, feat_dim: int = 64, classes: int = 19,
                        seed: int = 0, with_venues: bool = False):
  base = osp.join(path, size, 'processed')
  g = torch.Generator().manual_seed(seed)
  n = {'paper': papers, 'author': papers * 3 // 2, 'institute': max(papers // 40, 8), 'fos': max(papers // 20, 8)}
  topic_of = torch.randint(0, classes, (papers,), generator=g)
  proto = torch.randn(classes, feat_dim, generator=g)
  feats = {'paper': proto[topic_of] + 0.8 * torch.randn(papers, feat_dim, generator=g)}
  for nt in NTYPES[1:]:
    feats[nt] = torch.randn(n[nt], feat_dim, generator=g)

  def rnd(ns, nd, e):
    return torch.stack([torch.randint(0, ns, (e,), generator=g), torch.randint(0, nd, (e,), generator=g)], 1)
  # papers mostly cite papers of their own class (so that neighbours carry signal)
  src = torch.randint(0, papers, (papers * 6,), generator=g)
  cand = torch.randint(0, papers, (papers * 6,), generator=g)
  keep = (topic_of[src] == topic_of[cand]) | (torch.rand(papers * 6, generator=g) < 0.2)
  edges = {BASE_ETYPES[0]: torch.stack([src[keep], cand[keep]], 1),
           BASE_ETYPES[1]: rnd(papers, n['author'], papers * 3),
           BASE_ETYPES[2]: rnd(n['author'], n['institute'], n['author']),
           BASE_ETYPES[3]: rnd(papers, n['fos'], papers * 2)}
  ntypes = list(NTYPES)
  if with_venues:               # the schema of IGBH-large / -full: every paper has one journal or conference
    n['journal'], n['conference'] = max(papers // 100, 4), max(papers // 100, 4)
    for nt in VENUE_NTYPES:
      feats[nt] = torch.randn(n[nt], feat_dim, generator=g)
    edges[VENUE_ETYPES[0]] = rnd(papers, n['journal'], papers // 2)
    edges[VENUE_ETYPES[1]] = rnd(papers, n['conference'], papers // 2)
    ntypes += VENUE_NTYPES
  for et, ei in edges.items():
    os.makedirs(osp.join(base, etype_dir(et)), exist_ok=True)
    np.save(osp.join(base, etype_dir(et), 'edge_index.npy'), ei.numpy())
  for nt in ntypes:
    os.makedirs(osp.join(base, nt), exist_ok=True)
    np.save(osp.join(base, nt, 'node_feat.npy'), feats[nt].numpy())
  np.save(osp.join(base, 'paper', 'node_label_19.npy'), topic_of.numpy())
  np.save(osp.join(base, 'paper', 'node_label_2K.npy'), topic_of.numpy())
  return base


def float2half(base_path: str):
  """node_feat.npy -> node_feat_fp16.pt for every node type (halves the feature store)."""
  for nt in ntypes_on_disk(base_path):
    out = osp.join(base_path, nt, 'node_feat_fp16.pt')
    if not osp.exists(out):
      torch.save(torch.from_numpy(np.array(np.load(osp.join(base_path, nt, 'node_feat.npy'), mmap_mode='r'))).half(), out)


class IGBHeteroDataset(object):
  """Loads the processed IGBH directory into plain tensors:
    edge_dict  {etype: edge_index [2,E]}             (layout COO)
               {etype: (compressed_0, compressed_1)} (layout CSC / CSR, written by compress_graph.py)
    feat_dict  {ntype: [N, F] fp32 or fp16}, label [N_paper], train_idx / val_idx
  Reverse relations are added so that messages reach every node type (reference dataset.py:212-226)."""

  def __init__(self, path: str, dataset_size: str = 'tiny', in_memory: bool = True, use_label_2K: bool = False,
               with_edges: bool = True, layout: str = 'COO', use_fp16: bool = False):
    self.base_path = osp.join(path, dataset_size, 'processed')
    assert osp.isdir(self.base_path), f'{self.base_path} not found (make_synthetic_igbh() writes one)'
    self.in_memory, self.layout, self.use_fp16 = in_memory, layout.upper(), use_fp16
    self.ntypes = ntypes_on_disk(self.base_path)
    self.edge_dict, self.feat_dict = {}, {}
    if use_fp16:
      float2half(self.base_path)
    if with_edges:
      self._load_edges()
    self._load_features()
    lab = 'node_label_2K.npy' if use_label_2K else 'node_label_19.npy'
    self.label = torch.from_numpy(np.load(osp.join(self.base_path, 'paper', lab))).long()
    self.num_classes = int(self.label.max()) + 1
    self.train_idx = self._opt('train_idx.pt')
    self.val_idx = self._opt('val_idx.pt')

  def _opt(self, name):
    p = osp.join(self.base_path, name)
    return torch.load(p) if osp.exists(p) else None

  def _load_edges(self):
    mode = None if self.in_memory else 'r'
    if self.layout == 'COO':
      for et in etypes_on_disk(self.base_path):
        ei = torch.from_numpy(np.array(np.load(osp.join(self.base_path, etype_dir(et), 'edge_index.npy'), mmap_mode=mode))).t()
        if et[0] == et[2]:        # cites: make it symmetric like the reference (add reverse + dedup not needed)
          self.edge_dict[et] = torch.cat([ei, ei.flip(0)], 1).contiguous()
        else:
          self.edge_dict[et] = ei.contiguous()
          self.edge_dict[(et[2], 'rev_' + et[1], et[0])] = ei.flip(0).contiguous()
    else:
      d = osp.join(self.base_path, self.layout)
      assert osp.isdir(d), f'run compress_graph.py --layout {self.layout} first'
      for name in sorted(os.listdir(d)):
        et = tuple(name.split('__'))
        self.edge_dict[et] = (torch.load(osp.join(d, name, 'compressed_0.pt')),
                              torch.load(osp.join(d, name, 'compressed_1.pt')))
    self.etypes = list(self.edge_dict.keys())

  def _load_features(self):
    for nt in self.ntypes:
      if self.use_fp16:
        self.feat_dict[nt] = torch.load(osp.join(self.base_path, nt, 'node_feat_fp16.pt'))
      else:
        mode = None if self.in_memory else 'r'
        self.feat_dict[nt] = torch.from_numpy(np.array(np.load(osp.join(self.base_path, nt, 'node_feat.npy'), mmap_mode=mode)))

  @property
  def num_nodes(self) -> Dict[str, int]:
    return {nt: int(f.shape[0]) for nt, f in self.feat_dict.items()}
