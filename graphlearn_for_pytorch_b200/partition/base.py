"""Offline partitioning: partitioner base class, on-disk format, loaders.

Parity: reference python/partition/base.py:43-907.  Layout written under `output_dir`
(identical to the reference's):

  META                                   pickle {num_parts, data_cls, node_types, edge_types}
  node_pb.pt | node_pb/<ntype>.pt        partition book tensors
  edge_pb.pt | edge_pb/<etype>.pt
  part{i}/graph[/<etype>]/{rows,cols,eids,weights?}.pt
  part{i}/{node_feat,edge_feat}[/<type>]/{feats.pkl,ids.pkl,cache_feats.pt?,cache_ids.pt?}
  graph[/<etype>]/...                    optional full-topology cache (graph_caching)
"""
import os
import pickle
from abc import ABC, abstractmethod
from typing import Dict, List, Optional, Tuple, Union

import torch

from ..typing import (EdgeType, FeaturePartitionData, GraphPartitionData, NodeType, TensorDataType, as_str)
from ..utils.common import append_tensor_to_file, ensure_dir, load_and_concatenate_tensors
from ..utils.tensor import convert_to_tensor, id2idx
from .partition_book import GLTPartitionBook, PartitionBook


# ----------------------------------------------------------------------------- save helpers
def save_meta(output_dir: str, num_parts: int, data_cls: str = 'homo',
              node_types: Optional[List[NodeType]] = None, edge_types: Optional[List[EdgeType]] = None):
  meta = {'num_parts': num_parts, 'data_cls': data_cls, 'node_types': node_types, 'edge_types': edge_types}
  path = os.path.join(output_dir, 'META')
  tmp = f'{path}.tmp{os.getpid()}'
  with open(tmp, 'wb') as f:
    pickle.dump(meta, f, pickle.HIGHEST_PROTOCOL)
  os.replace(tmp, path)          # several ranks may write the same META into a shared directory


def _typed_path(root: str, stem: str, t=None, ext='.pt'):
  if t is not None:
    sub = os.path.join(root, stem)
    ensure_dir(sub)
    return os.path.join(sub, f'{as_str(t)}{ext}')
  return os.path.join(root, f'{stem}{ext}')


def _plain(pb):
  """Tensor books are written as PLAIN tensors: loadable with `weights_only=True` and by the reference's loader
  (same on-disk format); they are wrapped into GLTPartitionBook again on load."""
  return pb.as_subclass(torch.Tensor) if isinstance(pb, torch.Tensor) else pb


def _save_atomic(obj, path: str):
  tmp = f'{path}.tmp{os.getpid()}'
  torch.save(obj, tmp)
  os.replace(tmp, path)


def save_node_pb(output_dir: str, node_pb: PartitionBook, ntype: Optional[NodeType] = None):
  _save_atomic(_plain(node_pb), _typed_path(output_dir, 'node_pb', ntype))


def save_edge_pb(output_dir: str, edge_pb: PartitionBook, etype: Optional[EdgeType] = None):
  _save_atomic(_plain(edge_pb), _typed_path(output_dir, 'edge_pb', etype))


def _graph_dir(root: str, etype=None):
  d = os.path.join(root, 'graph')
  if etype is not None:
    d = os.path.join(d, as_str(etype))
  ensure_dir(d)
  return d


def _save_graph(d: str, g: GraphPartitionData):
  torch.save(g.edge_index[0], os.path.join(d, 'rows.pt'))
  torch.save(g.edge_index[1], os.path.join(d, 'cols.pt'))
  torch.save(g.eids, os.path.join(d, 'eids.pt'))
  if g.weights is not None:
    torch.save(g.weights, os.path.join(d, 'weights.pt'))


def save_graph_partition(output_dir: str, partition_idx: int, graph_partition: GraphPartitionData,
                         etype: Optional[EdgeType] = None):
  _save_graph(_graph_dir(os.path.join(output_dir, f'part{partition_idx}'), etype), graph_partition)


def save_graph_cache(output_dir: str, graph_partition_list: List[GraphPartitionData],
                     etype: Optional[EdgeType] = None, with_edge_feat: bool = False):
  """Store the *whole* topology once so every partition can sample locally (graph_caching)."""
  if not graph_partition_list:
    return
  rows = torch.cat([g.edge_index[0] for g in graph_partition_list])
  cols = torch.cat([g.edge_index[1] for g in graph_partition_list])
  eids = torch.cat([g.eids for g in graph_partition_list])
  w = torch.cat([g.weights for g in graph_partition_list]) if graph_partition_list[0].weights is not None else None
  _save_graph(_graph_dir(output_dir, etype), GraphPartitionData((rows, cols), eids, w))


def _feat_dir(output_dir, partition_idx, group, t=None):
  d = os.path.join(output_dir, f'part{partition_idx}', group)
  if t is not None:
    d = os.path.join(d, as_str(t))
  ensure_dir(d)
  return d


def save_feature_partition(output_dir: str, partition_idx: int, feature_partition: FeaturePartitionData,
                           group: str = 'node_feat', graph_type=None):
  d = _feat_dir(output_dir, partition_idx, group, graph_type)
  for name in ('feats.pkl', 'ids.pkl'):
    if os.path.exists(os.path.join(d, name)):
      os.remove(os.path.join(d, name))
  append_tensor_to_file(os.path.join(d, 'feats.pkl'), feature_partition.feats)
  append_tensor_to_file(os.path.join(d, 'ids.pkl'), feature_partition.ids)
  if feature_partition.cache_feats is not None:
    torch.save(feature_partition.cache_feats, os.path.join(d, 'cache_feats.pt'))
    torch.save(feature_partition.cache_ids, os.path.join(d, 'cache_ids.pt'))


def save_feature_partition_chunk(output_dir: str, partition_idx: int, feature_partition: FeaturePartitionData,
                                 group: str = 'node_feat', graph_type=None):
  """Append one chunk (feats, ids) to a partition's feature files."""
  d = _feat_dir(output_dir, partition_idx, group, graph_type)
  append_tensor_to_file(os.path.join(d, 'feats.pkl'), feature_partition.feats)
  append_tensor_to_file(os.path.join(d, 'ids.pkl'), feature_partition.ids)


def save_feature_partition_cache(output_dir: str, partition_idx: int, feature_partition: FeaturePartitionData,
                                 group: str = 'node_feat', graph_type=None):
  d = _feat_dir(output_dir, partition_idx, group, graph_type)
  if feature_partition.cache_feats is not None:
    torch.save(feature_partition.cache_feats, os.path.join(d, 'cache_feats.pt'))
    torch.save(feature_partition.cache_ids, os.path.join(d, 'cache_ids.pt'))


def _for_each_part(fn, num_parts: int):
  """Run fn(p) for every partition on a few threads: each partition owns its files, and the row gathers, the
  serialisation memcpys and the file writes all release the GIL, so the writers overlap."""
  workers = min(num_parts, max(1, (os.cpu_count() or 2) // 2), 8)
  if workers <= 1:
    for p in range(num_parts):
      fn(p)
    return
  from concurrent.futures import ThreadPoolExecutor
  with ThreadPoolExecutor(workers) as pool:
    for f in [pool.submit(fn, p) for p in range(num_parts)]:
      f.result()                      # re-raises a writer's exception


# ----------------------------------------------------------------------------- partitioner
class PartitionerBase(ABC):
  """Splits nodes, then edges (by source or destination owner), then features; writes the
  on-disk layout above.  Subclasses only decide node ownership (`_partition_node`) and the
  optional per-partition hot-feature cache (`_cache_node`)."""

  def __init__(self, output_dir: str, num_parts: int, num_nodes: Union[int, Dict[NodeType, int]],
               edge_index, node_feat=None, node_feat_dtype: torch.dtype = torch.float32,
               edge_feat=None, edge_feat_dtype: torch.dtype = torch.float32, edge_weights=None,
               edge_assign_strategy: str = 'by_src', chunk_size: int = 10000):
    self.output_dir = output_dir
    ensure_dir(output_dir)
    self.num_parts = num_parts
    assert num_parts > 1
    self.num_nodes = num_nodes
    self.edge_index = convert_to_tensor(edge_index, dtype=torch.int64)
    self.node_feat = convert_to_tensor(node_feat, dtype=node_feat_dtype)
    self.edge_feat = convert_to_tensor(edge_feat, dtype=edge_feat_dtype)
    self.edge_weights = convert_to_tensor(edge_weights, dtype=torch.float32)
    if isinstance(num_nodes, dict):
      self.data_cls = 'hetero'
      self.node_types = list(num_nodes.keys())
      self.edge_types = list(self.edge_index.keys())
      self.num_edges = {et: len(ei[0]) for et, ei in self.edge_index.items()}
    else:
      self.data_cls = 'homo'
      self.node_types = self.edge_types = None
      self.num_edges = len(self.edge_index[0])
    self.edge_assign_strategy = edge_assign_strategy.lower()
    assert self.edge_assign_strategy in ('by_src', 'by_dst')
    self.chunk_size = chunk_size

  # ---- accessors
  def get_edge_index(self, etype=None):
    ei = self.edge_index[etype] if self.data_cls == 'hetero' else self.edge_index
    return ei[0], ei[1]

  def get_node_feat(self, ntype=None):
    if self.node_feat is None:
      return None
    return self.node_feat.get(ntype) if self.data_cls == 'hetero' else self.node_feat

  def get_edge_feat(self, etype=None):
    if self.edge_feat is None:
      return None
    return self.edge_feat.get(etype) if self.data_cls == 'hetero' else self.edge_feat

  def get_edge_weights(self, etype=None):
    if self.edge_weights is None:
      return None
    return self.edge_weights.get(etype) if self.data_cls == 'hetero' else self.edge_weights

  def _num_nodes(self, ntype=None):
    return self.num_nodes[ntype] if self.data_cls == 'hetero' else self.num_nodes

  # ---- to be provided by subclasses
  @abstractmethod
  def _partition_node(self, ntype: Optional[NodeType] = None) -> Tuple[List[torch.Tensor], PartitionBook]:
    """Returns (ids owned by each partition, node partition book)."""

  def _cache_node(self, ntype: Optional[NodeType] = None) -> List[Optional[torch.Tensor]]:
    """Ids whose features each partition additionally caches (hot remote rows)."""
    return [None] * self.num_parts

  # ---- graph
  def _partition_graph(self, node_pbs, etype: Optional[EdgeType] = None):
    rows, cols = self.get_edge_index(etype)
    weights = self.get_edge_weights(etype)
    n_e = rows.numel()
    eids = torch.arange(n_e, dtype=torch.int64)
    if self.data_cls == 'hetero':
      pb = node_pbs[etype[0]] if self.edge_assign_strategy == 'by_src' else node_pbs[etype[2]]
    else:
      pb = node_pbs
    key = rows if self.edge_assign_strategy == 'by_src' else cols
    edge_pb = torch.empty(n_e, dtype=torch.int64)
    chunk = max(self.chunk_size, 1) * 64
    for b in range(0, n_e, chunk):
      edge_pb[b:b + chunk] = pb[key[b:b + chunk]]
    parts = []
    for p in range(self.num_parts):
      m = edge_pb == p
      parts.append(GraphPartitionData((rows[m], cols[m]), eids[m], weights[m] if weights is not None else None))
    return parts, GLTPartitionBook(edge_pb)

  # ---- features
  def _save_feats(self, feat, ids_per_part, cache_ids, group, t):
    def save_part(p):
      d = _feat_dir(self.output_dir, p, group, t)
      for name in ('feats.pkl', 'ids.pkl'):
        if os.path.exists(os.path.join(d, name)):
          os.remove(os.path.join(d, name))
      ids = ids_per_part[p]
      step = max(self.chunk_size, 1)
      if ids.numel() == 0:
        save_feature_partition_chunk(self.output_dir, p, FeaturePartitionData(feat[ids], ids), group, t)
      for b in range(0, ids.numel(), step):
        c = ids[b:b + step]
        save_feature_partition_chunk(self.output_dir, p, FeaturePartitionData(feat[c], c), group, t)
      if cache_ids is not None and cache_ids[p] is not None and cache_ids[p].numel() > 0:
        save_feature_partition_cache(self.output_dir, p,
                                     FeaturePartitionData(None, None, feat[cache_ids[p]], cache_ids[p]), group, t)
    _for_each_part(save_part, self.num_parts)

  def _process_node(self, ntype, with_feature):
    ids_per_part, pb = self._partition_node(ntype)
    save_node_pb(self.output_dir, pb, ntype)
    feat = self.get_node_feat(ntype)
    if with_feature and feat is not None:
      self._save_feats(feat, ids_per_part, self._cache_node(ntype), 'node_feat', ntype)
    return pb

  def _process_edge(self, node_pbs, etype, with_feature, graph_caching):
    parts, edge_pb = self._partition_graph(node_pbs, etype)
    save_edge_pb(self.output_dir, edge_pb, etype)
    if graph_caching:
      save_graph_cache(self.output_dir, parts, etype)
    else:
      _for_each_part(lambda p: save_graph_partition(self.output_dir, p, parts[p], etype), self.num_parts)
    feat = self.get_edge_feat(etype)
    if with_feature and feat is not None:
      self._save_feats(feat, [g.eids for g in parts], None, 'edge_feat', etype)

  def partition(self, with_feature: bool = True, graph_caching: bool = False):
    """Run the whole pipeline and write `output_dir`."""
    if self.data_cls == 'hetero':
      node_pbs = {nt: self._process_node(nt, with_feature) for nt in self.node_types}
      for et in self.edge_types:
        self._process_edge(node_pbs, et, with_feature, graph_caching)
    else:
      node_pb = self._process_node(None, with_feature)
      self._process_edge(node_pb, None, with_feature, graph_caching)
    save_meta(self.output_dir, self.num_parts, self.data_cls, self.node_types, self.edge_types)


# ----------------------------------------------------------------------------- loaders
def _load_graph_dir(d: str, device) -> Optional[GraphPartitionData]:
  if not os.path.exists(os.path.join(d, 'rows.pt')):
    return None
  rows = torch.load(os.path.join(d, 'rows.pt'), map_location=device)
  cols = torch.load(os.path.join(d, 'cols.pt'), map_location=device)
  # eids.pt is optional (the reference's save_graph_cache only writes it with edge features,
  # python/partition/base.py): absent = edges are numbered in file order
  ep = os.path.join(d, 'eids.pt')
  eids = torch.load(ep, map_location=device) if os.path.exists(ep) else \
      torch.arange(rows.numel(), dtype=torch.int64, device=rows.device)
  wp = os.path.join(d, 'weights.pt')
  w = torch.load(wp, map_location=device) if os.path.exists(wp) else None
  return GraphPartitionData((rows, cols), eids, w)


def load_graph_partition_data(graph_data_dir: str, device: torch.device) -> Optional[GraphPartitionData]:
  return _load_graph_dir(graph_data_dir, device)


def load_feature_partition_data(feature_data_dir: str, device: torch.device) -> Optional[FeaturePartitionData]:
  fp, ip = os.path.join(feature_data_dir, 'feats.pkl'), os.path.join(feature_data_dir, 'ids.pkl')
  if not (os.path.exists(fp) and os.path.exists(ip)):
    return None
  feats = load_and_concatenate_tensors(fp, device)
  ids = load_and_concatenate_tensors(ip, device)
  cf, ci = os.path.join(feature_data_dir, 'cache_feats.pt'), os.path.join(feature_data_dir, 'cache_ids.pt')
  cache_feats = torch.load(cf, map_location=device) if os.path.exists(cf) else None
  cache_ids = torch.load(ci, map_location=device) if os.path.exists(ci) else None
  return FeaturePartitionData(feats, ids, cache_feats, cache_ids)


def _load_pb(path, device):
  try:
    pb = torch.load(path, map_location=device, weights_only=True)
  except Exception:   # range books (objects) and files written before tensor books were stored plain
    pb = torch.load(path, map_location=device, weights_only=False)
  if type(pb) is torch.Tensor:
    pb = GLTPartitionBook(pb)
  return pb


def load_partition(root_dir: str, partition_idx: int, graph_caching: bool = False,
                   device: torch.device = torch.device('cpu')):
  """-> (num_parts, partition_idx, graph_data, node_feat_data, edge_feat_data, node_pb, edge_pb);
  dicts keyed by type for hetero data."""
  with open(os.path.join(root_dir, 'META'), 'rb') as f:
    meta = pickle.load(f)
  num_parts = meta['num_parts']
  assert 0 <= partition_idx < num_parts
  part_dir = os.path.join(root_dir, f'part{partition_idx}')
  graph_root = os.path.join(root_dir, 'graph') if graph_caching else os.path.join(part_dir, 'graph')
  if meta['data_cls'] == 'homo':
    graph = load_graph_partition_data(graph_root, device)
    node_feat = load_feature_partition_data(os.path.join(part_dir, 'node_feat'), device)
    edge_feat = load_feature_partition_data(os.path.join(part_dir, 'edge_feat'), device)
    node_pb = _load_pb(os.path.join(root_dir, 'node_pb.pt'), device)
    edge_pb = _load_pb(os.path.join(root_dir, 'edge_pb.pt'), device)
    return num_parts, partition_idx, graph, node_feat, edge_feat, node_pb, edge_pb
  graph, node_feat, edge_feat, node_pb, edge_pb = {}, {}, {}, {}, {}
  for et in meta['edge_types']:
    g = load_graph_partition_data(os.path.join(graph_root, as_str(et)), device)
    if g is not None:
      graph[tuple(et)] = g
    ef = load_feature_partition_data(os.path.join(part_dir, 'edge_feat', as_str(et)), device)
    if ef is not None:
      edge_feat[tuple(et)] = ef
    edge_pb[tuple(et)] = _load_pb(os.path.join(root_dir, 'edge_pb', f'{as_str(et)}.pt'), device)
  for nt in meta['node_types']:
    nf = load_feature_partition_data(os.path.join(part_dir, 'node_feat', as_str(nt)), device)
    if nf is not None:
      node_feat[nt] = nf
    node_pb[nt] = _load_pb(os.path.join(root_dir, 'node_pb', f'{as_str(nt)}.pt'), device)
  return (num_parts, partition_idx, graph, node_feat or None, edge_feat or None, node_pb, edge_pb)


def cat_feature_cache(partition_idx: int, feat_pdata: FeaturePartitionData, feat_pb: PartitionBook):
  """Prepend the cached hot rows of remote ids to the local rows and rewrite the feature
  partition book so those ids resolve locally.  -> (cache_ratio, feats, id2index, new_pb)."""
  feats, ids = feat_pdata.feats, feat_pdata.ids
  cache_feats, cache_ids = feat_pdata.cache_feats, feat_pdata.cache_ids
  if cache_feats is None or cache_ids is None:
    return 0.0, feats, id2idx(ids), feat_pb
  ratio = cache_ids.size(0) / (cache_ids.size(0) + ids.size(0))
  new_feats = torch.cat([cache_feats, feats])
  max_id = int(max(cache_ids.max().item(), ids.max().item()))
  nid2idx = torch.zeros(max_id + 1, dtype=torch.int64, device=feats.device)
  nid2idx[ids] = torch.arange(ids.size(0), dtype=torch.int64, device=feats.device) + cache_ids.size(0)
  nid2idx[cache_ids] = torch.arange(cache_ids.size(0), dtype=torch.int64, device=feats.device)
  new_pb = feat_pb.clone()
  new_pb[cache_ids] = partition_idx
  return ratio, new_feats, nid2idx, new_pb


def build_partition_feature(root_dir: str, partition_idx: int, chunk_size: int = 10000,
                            node_feat=None, node_feat_dtype: torch.dtype = torch.float32,
                            edge_feat=None, edge_feat_dtype: torch.dtype = torch.float32):
  """Two-stage partitioning: the topology was partitioned earlier (with_feature=False);
  now write the feature files of ONE partition from the full feature tensors, using the
  stored partition books (reference base.py:585-702)."""
  with open(os.path.join(root_dir, 'META'), 'rb') as f:
    meta = pickle.load(f)
  node_feat = convert_to_tensor(node_feat, dtype=node_feat_dtype)
  edge_feat = convert_to_tensor(edge_feat, dtype=edge_feat_dtype)

  def _write(feat, ids, group, t):
    d = _feat_dir(root_dir, partition_idx, group, t)
    for name in ('feats.pkl', 'ids.pkl'):
      if os.path.exists(os.path.join(d, name)):
        os.remove(os.path.join(d, name))
    for b in range(0, max(ids.numel(), 1), max(chunk_size, 1)):
      c = ids[b:b + chunk_size]
      save_feature_partition_chunk(root_dir, partition_idx, FeaturePartitionData(feat[c], c), group, t)

  if meta['data_cls'] == 'homo':
    if node_feat is not None:
      pb = _load_pb(os.path.join(root_dir, 'node_pb.pt'), 'cpu')
      _write(node_feat, torch.where(pb == partition_idx)[0], 'node_feat', None)
    if edge_feat is not None:
      g = load_graph_partition_data(os.path.join(root_dir, f'part{partition_idx}', 'graph'), 'cpu')
      if g is not None:
        owned = g.eids
      else:
        # topology written with graph_caching=True (no per-partition graph dir): edge ownership
        # comes from the stored edge partition book
        epb = _load_pb(os.path.join(root_dir, 'edge_pb.pt'), 'cpu')
        owned = torch.where(torch.as_tensor(epb) == partition_idx)[0]
      _write(edge_feat, owned, 'edge_feat', None)
    return
  for nt in meta['node_types']:
    if node_feat is not None and nt in node_feat:
      pb = _load_pb(os.path.join(root_dir, 'node_pb', f'{as_str(nt)}.pt'), 'cpu')
      _write(node_feat[nt], torch.where(pb == partition_idx)[0], 'node_feat', nt)
  for et in meta['edge_types']:
    et = tuple(et)
    if edge_feat is not None and et in edge_feat:
      g = load_graph_partition_data(os.path.join(root_dir, f'part{partition_idx}', 'graph', as_str(et)), 'cpu')
      if g is not None:
        _write(edge_feat[et], g.eids, 'edge_feat', et)
      else:
        epb_path = os.path.join(root_dir, 'edge_pb', f'{as_str(et)}.pt')
        if os.path.exists(epb_path):
          epb = _load_pb(epb_path, 'cpu')
          _write(edge_feat[et], torch.where(torch.as_tensor(epb) == partition_idx)[0], 'edge_feat', et)
