"""DistRandomPartitioner: partition a dataset that is *already spread* over several processes.

Parity: reference python/distributed/dist_random_partitioner.py:36-539.  Each rank holds a
slice of the node ids / edges / feature rows.  Node ownership is drawn at random per slice;
the slices of the partition book are all-gathered; every rank then ships each edge /
feature row to its owner over RPC (`DistPartitionManager`) and keeps what it receives.
The result equals the on-disk layout of `partition.RandomPartitioner`, one part per rank.
"""
import threading
from typing import Dict, List, Optional, Tuple, Union

import torch

from ..partition import (GLTPartitionBook, PartitionBook, save_edge_pb, save_feature_partition,
                         save_graph_partition, save_meta, save_node_pb)
from ..typing import EdgeType, FeaturePartitionData, GraphPartitionData, NodeType, as_str
from ..utils.common import ensure_dir
from ..utils.tensor import convert_to_tensor
from .dist_context import get_context, init_worker_group
from .rpc import (RpcCalleeBase, all_gather, barrier, get_rpc_current_group_worker_names, init_rpc,
                  rpc_is_initialized, rpc_register, rpc_request_async)


class DistPartitionManager(object):
  """Receives pieces (dicts of tensors tagged by a key) from every rank and concatenates."""

  def __init__(self):
    self._lock = threading.Lock()
    self._store: Dict[str, List[Dict[str, torch.Tensor]]] = {}

  def add(self, key: str, piece: Dict[str, torch.Tensor]):
    with self._lock:
      self._store.setdefault(key, []).append(piece)
    return True

  def collect(self, key: str) -> Dict[str, torch.Tensor]:
    with self._lock:
      pieces = self._store.pop(key, [])
    if not pieces:
      return {}
    return {k: torch.cat([p[k] for p in pieces]) for k in pieces[0].keys()}


class _PartitionCallee(RpcCalleeBase):
  def __init__(self, mgr: DistPartitionManager):
    self.mgr = mgr

  def call(self, key, piece):
    return self.mgr.add(key, piece)


class DistRandomPartitioner(object):
  """Args:
    output_dir: where this rank writes its `part{rank}` (and rank 0 the books/META).
    num_nodes: total node count (dict for hetero).
    edge_index: this rank's slice of the edges ([2, e] or dict).
    edge_ids: global ids of those edges.
    node_feat / node_feat_ids: this rank's slice of node features and their global ids.
    edge_feat / edge_feat_ids: same for edges.
    num_parts / current_partition_idx: default to the worker group size / rank.
    edge_assign_strategy: 'by_src' | 'by_dst'.
  """

  def __init__(self, output_dir: str, num_nodes, edge_index, edge_ids, node_feat=None, node_feat_ids=None,
               edge_feat=None, edge_feat_ids=None, num_parts: Optional[int] = None,
               current_partition_idx: Optional[int] = None, node_feat_dtype=torch.float32,
               edge_feat_dtype=torch.float32, edge_assign_strategy: str = 'by_src', chunk_size: int = 10000,
               master_addr: Optional[str] = None, master_port: Optional[int] = None,
               num_rpc_threads: int = 16):
    self.output_dir = output_dir
    ensure_dir(output_dir)
    ctx = get_context()
    if ctx is not None:
      self.num_parts, self.rank = ctx.world_size, ctx.rank
    else:
      assert num_parts is not None and current_partition_idx is not None
      self.num_parts, self.rank = num_parts, current_partition_idx
      init_worker_group(num_parts, current_partition_idx, 'dist_random_partitioner')
    if num_parts is not None:
      assert num_parts == self.num_parts
    if not rpc_is_initialized():
      assert master_addr is not None and master_port is not None
      init_rpc(master_addr, master_port, num_rpc_threads)
    self.num_nodes = num_nodes
    self.edge_index = convert_to_tensor(edge_index, dtype=torch.int64)
    self.edge_ids = convert_to_tensor(edge_ids, dtype=torch.int64)
    self.node_feat = convert_to_tensor(node_feat, dtype=node_feat_dtype)
    self.node_feat_ids = convert_to_tensor(node_feat_ids, dtype=torch.int64)
    self.edge_feat = convert_to_tensor(edge_feat, dtype=edge_feat_dtype)
    self.edge_feat_ids = convert_to_tensor(edge_feat_ids, dtype=torch.int64)
    self.data_cls = 'hetero' if isinstance(num_nodes, dict) else 'homo'
    self.node_types = list(num_nodes.keys()) if self.data_cls == 'hetero' else None
    self.edge_types = list(self.edge_index.keys()) if self.data_cls == 'hetero' else None
    self.edge_assign_strategy = edge_assign_strategy.lower()
    self.chunk_size = chunk_size
    self._mgr = DistPartitionManager()
    self._callee_id = rpc_register(_PartitionCallee(self._mgr))
    self._workers = get_rpc_current_group_worker_names()

  # ---- helpers
  def _ship(self, key: str, owners: torch.Tensor, tensors: Dict[str, torch.Tensor]):
    """Send rows of `tensors` to their owner ranks; returns what this rank ends up with.

    The slice is walked in chunks of `chunk_size` rows (reference python/distributed/dist_random_partitioner.py:
    257-290, `_partition_by_chunk`): every chunk is bucketed by owner and shipped as one RPC per destination, with
    at most `max_inflight` requests outstanding, so the peak extra memory is O(chunk_size) per destination instead
    of a second copy of the whole slice (IGBH-scale inputs do not fit twice) and a slow receiver back-pressures
    the sender instead of queueing everything."""
    n = owners.numel()
    chunk = max(int(self.chunk_size), 1)
    max_inflight = max(2 * self.num_parts, 8)
    futs = []
    for b in range(0, max(n, 1), chunk):
      own = owners[b:b + chunk]
      if own.numel() == 0:
        break
      # one stable sort per chunk instead of num_parts boolean masks
      order = torch.argsort(own.to(torch.int16), stable=True)   # narrow keys take torch's radix path
      counts = torch.bincount(own, minlength=self.num_parts).tolist()
      sorted_chunk = {k: v[b:b + chunk][order] for k, v in tensors.items()}
      off = 0
      for p_ in range(self.num_parts):
        c = counts[p_]
        if c == 0:
          continue
        piece = {k: v[off:off + c].clone() for k, v in sorted_chunk.items()}
        off += c
        if p_ == self.rank:
          self._mgr.add(key, piece)
        else:
          futs.append(rpc_request_async(self._workers[p_], self._callee_id, args=(key, piece)))
          if len(futs) >= max_inflight:
            futs.pop(0).wait()
    for f in futs:
      f.wait()
    barrier()
    out = self._mgr.collect(key)
    barrier()
    return out

  def _node_book(self, ntype=None) -> torch.Tensor:
    n = self.num_nodes[ntype] if ntype is not None else self.num_nodes
    per = (n + self.num_parts - 1) // self.num_parts
    lo, hi = min(self.rank * per, n), min((self.rank + 1) * per, n)
    g = torch.Generator()
    g.manual_seed(1234 + self.rank + (hash(ntype) % 1000 if ntype else 0))
    # balanced: a random permutation of this slice dealt round-robin over the partitions (the remainder of each
    # slice goes to different partitions on different ranks), so every partition gets n / num_parts nodes like the
    # reference's chunked assignment (dist_random_partitioner.py:292-318)
    local = torch.empty(hi - lo, dtype=torch.int64)
    local[torch.randperm(hi - lo, generator=g)] = (torch.arange(hi - lo) + self.rank) % self.num_parts
    gathered = all_gather((lo, local))
    book = torch.empty(n, dtype=torch.int64)
    for name in self._workers:
      l, part = gathered[name]
      book[l:l + part.numel()] = part
    return book

  def _partition_one_graph(self, node_pbs, etype=None):
    ei = self.edge_index[etype] if etype is not None else self.edge_index
    eids = self.edge_ids[etype] if etype is not None else self.edge_ids
    if etype is not None:
      pb = node_pbs[etype[0]] if self.edge_assign_strategy == 'by_src' else node_pbs[etype[2]]
    else:
      pb = node_pbs
    owners = pb[ei[0]] if self.edge_assign_strategy == 'by_src' else pb[ei[1]]
    key = f'graph:{as_str(etype) if etype else ""}'
    got = self._ship(key, owners, {'rows': ei[0], 'cols': ei[1], 'eids': eids})
    # edge partition book: every rank contributes (eid, owner) of its slice
    gathered = all_gather((eids, owners))
    n_e = sum(v[0].numel() for v in gathered.values())
    max_e = max([int(v[0].max()) + 1 if v[0].numel() else 0 for v in gathered.values()] + [n_e])
    edge_pb = torch.zeros(max_e, dtype=torch.int64)
    for v in gathered.values():
      edge_pb[v[0]] = v[1]
    e = torch.empty(0, dtype=torch.int64)
    return GraphPartitionData((got.get('rows', e), got.get('cols', e)), got.get('eids', e)), edge_pb

  def _partition_one_feat(self, feat, ids, pb, key):
    if feat is None:
      return None
    got = self._ship(key, pb[ids], {'feats': feat, 'ids': ids})
    return FeaturePartitionData(got['feats'], got['ids'])

  # ---- driver
  def partition(self):
    """Run the collective partitioning and write this rank's part to `output_dir`."""
    if self.data_cls == 'hetero':
      node_pbs = {nt: self._node_book(nt) for nt in self.node_types}
      for nt in self.node_types:
        save_node_pb(self.output_dir, GLTPartitionBook(node_pbs[nt]), nt)     # every rank: output dirs may differ
        if self.node_feat is not None and nt in self.node_feat:
          f = self._partition_one_feat(self.node_feat[nt], self.node_feat_ids[nt], node_pbs[nt], f'nfeat:{nt}')
          save_feature_partition(self.output_dir, self.rank, f, 'node_feat', nt)
      for et in self.edge_types:
        g, epb = self._partition_one_graph(node_pbs, et)
        save_graph_partition(self.output_dir, self.rank, g, et)
        save_edge_pb(self.output_dir, GLTPartitionBook(epb), et)
        if self.edge_feat is not None and et in self.edge_feat:
          f = self._partition_one_feat(self.edge_feat[et], self.edge_feat_ids[et], epb, f'efeat:{as_str(et)}')
          save_feature_partition(self.output_dir, self.rank, f, 'edge_feat', et)
    else:
      node_pb = self._node_book()
      save_node_pb(self.output_dir, GLTPartitionBook(node_pb))
      if self.node_feat is not None:
        f = self._partition_one_feat(self.node_feat, self.node_feat_ids, node_pb, 'nfeat')
        save_feature_partition(self.output_dir, self.rank, f, 'node_feat')
      g, epb = self._partition_one_graph(node_pb)
      save_graph_partition(self.output_dir, self.rank, g)
      save_edge_pb(self.output_dir, GLTPartitionBook(epb))
      if self.edge_feat is not None:
        f = self._partition_one_feat(self.edge_feat, self.edge_feat_ids, epb, 'efeat')
        save_feature_partition(self.output_dir, self.rank, f, 'edge_feat')
    save_meta(self.output_dir, self.num_parts, self.data_cls, self.node_types, self.edge_types)
    barrier()
