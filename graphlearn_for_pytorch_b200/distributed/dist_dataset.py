"""DistDataset: one partition of a partitioned dataset + the books to find the rest.

Parity: reference python/distributed/dist_dataset.py:30-318 (load from the on-disk layout,
feature-cache concatenation, random split over owned ids, IPC pickling).  Extra:
`DistDataset.from_p2p(...)` wraps NVLink-mapped multi-shard tables (parallel/) so the same
loaders run with in-kernel peer reads instead of RPC.
"""
from multiprocessing.reduction import ForkingPickler
from typing import Dict, List, Optional, Union

import torch

from ..data import Dataset, DeviceGroup, Feature, Graph
from ..partition import PartitionBook, RangePartitionBook, cat_feature_cache, load_partition
from ..typing import EdgeType, FeaturePartitionData, GraphPartitionData, NodeType
from ..utils.common import default_id_filter, default_id_select
from ..utils.tensor import convert_to_tensor, id2idx, share_memory, squeeze


class DistDataset(Dataset):
  """One partition of a partitioned dataset + the partition books that route ids to their owners.

  `load(root, partition_idx)` reads the on-disk format written by the partitioners (feature caches are prepended and
  the feature book rewritten so cached remote rows resolve locally); `load_vineyard` builds it from a property-graph
  fragment; `from_p2p` wraps NVLink-mapped shards (no RPC on the data path).  (Reference:
  python/distributed/dist_dataset.py:30-330.)"""
  def __init__(self, num_partitions: int = 1, partition_idx: int = 0, graph_partition=None,
               node_feature_partition=None, edge_feature_partition=None, whole_node_labels=None,
               node_pb=None, edge_pb=None, node_feat_pb=None, edge_feat_pb=None, edge_dir: str = 'out',
               graph_caching: bool = False, node_split=None, id_filter=default_id_filter,
               id_select=default_id_select):
    """graph_caching: the partition directory holds the FULL topology (`load(..., graph_caching=True)`).
    id_filter(node_pb, partition_idx) -> ids owned by this partition; id_select(ids, mask, node_pb) -> the ids of a
    request that a given partition serves: hooks for partition books that are not plain tensors (the vineyard bridge
    installs its own; reference dist_dataset.py:47-64)."""
    super().__init__(graph_partition, node_feature_partition, edge_feature_partition, whole_node_labels,
                     edge_dir, node_split)
    self.graph_caching = graph_caching
    self.id_filter = id_filter
    self.id_select = id_select
    self.num_partitions = num_partitions
    self.partition_idx = partition_idx
    self.node_pb = node_pb
    self.edge_pb = edge_pb
    # after cat_feature_cache the *feature* books differ from the graph books
    self._node_feat_pb = node_feat_pb
    self._edge_feat_pb = edge_feat_pb
    self.data_plane = 'rpc'
    self.p2p_bounds = None
    if self.graph is not None:
      assert self.node_pb is not None

  # ------------------------------------------------------------------ load
  def load(self, root_dir: str, partition_idx: int, graph_mode: str = 'ZERO_COPY', input_layout: str = 'COO',
           feature_with_gpu: bool = True, graph_caching: bool = False,
           device_group_list: Optional[List[DeviceGroup]] = None,
           whole_node_label_file: Union[str, Dict[NodeType, str], None] = None,
           device: Optional[int] = None):
    self.graph_caching = graph_caching
    (self.num_partitions, self.partition_idx, graph_data, node_feat_data, edge_feat_data, node_pb,
     edge_pb) = load_partition(root_dir, partition_idx, graph_caching)
    if isinstance(graph_data, dict):
      ei, eids, w = {}, {}, {}
      for et, g in graph_data.items():
        ei[et] = torch.stack(list(g.edge_index))
        eids[et] = g.eids
        if g.weights is not None:
          w[et] = g.weights
      self.init_graph(ei, eids, w or None, layout=input_layout, graph_mode=graph_mode, device=device)
    else:
      self.init_graph(torch.stack(list(graph_data.edge_index)), graph_data.eids, graph_data.weights,
                      layout=input_layout, graph_mode=graph_mode, device=device)
    self.node_pb, self.edge_pb = node_pb, edge_pb

    def build(feat_data, pb, is_node):
      if feat_data is None:
        return None, None
      if isinstance(feat_data, dict):
        feats, ids2idx, ratios, books = {}, {}, {}, {}
        for t, fd in feat_data.items():
          ratios[t], feats[t], ids2idx[t], books[t] = cat_feature_cache(partition_idx, fd, pb[t])
        init = self.init_node_features if is_node else self.init_edge_features
        init(feats, ids2idx, split_ratio=ratios, device_group_list=device_group_list, device=device,
             with_gpu=feature_with_gpu, dtype=next(iter(feats.values())).dtype)
        return books, None
      ratio, feats, i2i, book = cat_feature_cache(partition_idx, feat_data, pb)
      init = self.init_node_features if is_node else self.init_edge_features
      init(feats, i2i, split_ratio=ratio, device_group_list=device_group_list, device=device,
           with_gpu=feature_with_gpu, dtype=feats.dtype)
      return book, None
    self._node_feat_pb, _ = build(node_feat_data, node_pb, True)
    self._edge_feat_pb, _ = build(edge_feat_data, edge_pb, False)
    if whole_node_label_file is not None:
      if isinstance(whole_node_label_file, dict):
        self.init_node_labels({nt: torch.load(f) for nt, f in whole_node_label_file.items()})
      else:
        self.init_node_labels(torch.load(whole_node_label_file))
    return self

  def load_vineyard(self, vineyard_id: str, vineyard_socket: str, edges, edge_weights=None, node_features=None,
                    edge_features=None, node_labels=None, graph_mode: str = 'CPU', feature_with_gpu: bool = False,
                    fid2pid: Optional[Dict[int, int]] = None):
    """One property-graph fragment per worker (reference dist_dataset.py:215-243): partition id =
    fragment id, node partition books resolve gid -> fragment through the fragment store.  The
    topology and labels are keyed by global id so that the RPC sampler / feature lookups address
    them exactly like the partitions produced by the partitioners."""
    from ..data import vineyard_utils as vu
    frag = vu._open(vineyard_socket, vineyard_id)
    self.num_partitions, self.partition_idx = frag.fnum, frag.fid
    if fid2pid is not None:
      self.partition_idx = int(fid2pid[frag.fid])
    super().load_vineyard(vineyard_id, vineyard_socket, edges, edge_weights, node_features, edge_features,
                          node_labels, graph_mode=graph_mode, with_gpu=feature_with_gpu, global_rows=True)
    is_homo = len(edges) == 1 and edges[0][0] == edges[0][2]
    ntypes = sorted({et[0] for et in edges} | {et[2] for et in edges})
    books = {nt: vu.VineyardPartitionBook(vineyard_socket, vineyard_id, nt, fid2pid) for nt in ntypes}
    self.node_pb = books[ntypes[0]] if is_homo else books
    self.edge_pb = None
    if node_features:
      self._node_feat_pb = self.node_pb if is_homo else {nt: books[nt] for nt in node_features}
    self.id_select = vu.v6d_id_select
    self.id_filter = vu.v6d_id_filter
    return self

  # ------------------------------------------------------------------ p2p
  @classmethod
  def from_p2p(cls, partitioned_graph, partitioned_feature=None, labels=None, edge_dir: str = 'out'):
    """Wrap NVLink-mapped tables: every rank can read every shard in-kernel, so samplers and
    feature lookups built on this dataset never issue RPCs."""
    ranges = [(partitioned_graph.bounds[r], partitioned_graph.bounds[r + 1]) for r in range(partitioned_graph.world)]
    pb = RangePartitionBook(ranges, partitioned_graph.rank)
    ds = cls(partitioned_graph.world, partitioned_graph.rank, partitioned_graph.graph, None, None, labels, pb, None,
             pb, None, edge_dir)
    ds.node_features = partitioned_feature
    ds.data_plane = 'p2p'
    ds.p2p_bounds = list(partitioned_graph.bounds)
    return ds

  # ------------------------------------------------------------------ splits
  def random_node_split(self, num_val: Union[float, int], num_test: Union[float, int]):
    """Split the ids *owned by this partition* into train/val/test."""
    from ..data.dataset import random_split

    def owned(pb, n_hint=None):
      if hasattr(pb, 'id_filter'):
        return pb.id_filter(pb, self.partition_idx)
      return self.id_filter(pb, self.partition_idx)
    if isinstance(self.node_pb, dict):
      tr, va, te = {}, {}, {}
      for nt, pb in self.node_pb.items():
        ids = owned(pb)
        a, b, c = random_split(ids.numel(), num_val, num_test)
        tr[nt], va[nt], te[nt] = ids[a], ids[b], ids[c]
      self.train_idx, self.val_idx, self.test_idx = tr, va, te
    else:
      ids = owned(self.node_pb)
      a, b, c = random_split(ids.numel(), num_val, num_test)
      self.train_idx, self.val_idx, self.test_idx = ids[a], ids[b], ids[c]

  # ------------------------------------------------------------------ books
  @property
  def node_feat_pb(self):
    return self._node_feat_pb if self._node_feat_pb is not None else self.node_pb

  @property
  def edge_feat_pb(self):
    return self._edge_feat_pb if self._edge_feat_pb is not None else self.edge_pb

  # ------------------------------------------------------------------ IPC
  def share_ipc(self):
    base = super().share_ipc()
    for pb in (self.node_pb, self.edge_pb, self._node_feat_pb, self._edge_feat_pb):
      share_memory(pb if not isinstance(pb, dict) else list(pb.values()))
    return (self.num_partitions, self.partition_idx, base, self.node_pb, self.edge_pb,
            self._node_feat_pb, self._edge_feat_pb)

  @classmethod
  def from_ipc_handle(cls, ipc_handle):
    (num_partitions, partition_idx, base, node_pb, edge_pb, nfpb, efpb) = ipc_handle
    g, nf, ef, nl, edge_dir, split = base
    return cls(num_partitions, partition_idx, g, nf, ef, nl, node_pb, edge_pb, nfpb, efpb, edge_dir, node_split=split)


def rebuild_dist_dataset(ipc_handle):
  return DistDataset.from_ipc_handle(ipc_handle)


def reduce_dist_dataset(ds: DistDataset):
  return (rebuild_dist_dataset, (ds.share_ipc(),))


ForkingPickler.register(DistDataset, reduce_dist_dataset)
