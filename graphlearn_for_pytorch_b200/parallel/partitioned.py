"""Range-partitioned graph / feature store over the GPUs of one NVSwitch box.

Rank r owns rows [bounds[r], bounds[r+1]) of the CSR and of every feature table; all
ranks hold a GraphHandle / RowTableHandle whose shard pointers cover *all* ranks, so the
sampling and gather kernels resolve owner(v) in-kernel and read the owner's HBM
directly.  This replaces the reference's partition-book lookup + RPC fan-out + stitch
(distributed/dist_neighbor_sampler.py:616-687, distributed/dist_feature.py:176-452).
"""
from typing import List, Optional

import torch

from ..data.graph import Graph, Topology
from ..data.unified_tensor import UnifiedTensor
from ..ops import require_native
from .peer import exchange_peer_tensors, range_bounds, world_info


def shard_topology(topo: Topology, bounds: List[int], rank: int, device: torch.device,
                   idx_dtype=None):
  """Slice rank's row range out of a full CSR and move it to `device`."""
  b, e = bounds[rank], bounds[rank + 1]
  indptr = topo.indptr
  lo, hi = int(indptr[b]), int(indptr[e])
  max_id = int(topo.indices.max()) if topo.indices.numel() else 0
  if idx_dtype is None:
    idx_dtype = torch.int32 if max_id < 2 ** 31 - 1 else torch.int64
  out = {
    'indptr': (indptr[b:e + 1] - lo).to(device).contiguous(),
    'indices': topo.indices[lo:hi].to(device, dtype=idx_dtype).contiguous(),
    'eids': topo.edge_ids[lo:hi].to(device).contiguous() if topo.edge_ids is not None else None,
    'weights': topo.edge_weights[lo:hi].to(device).contiguous() if topo.edge_weights is not None else None,
    'row_begin': b, 'row_end': e,
  }
  return out


class PartitionedGraph(object):
  """Collectively build a multi-shard `Graph` from each rank's local shard."""

  def __init__(self, local_shard: dict, bounds: List[int], device: torch.device, group=None):
    rank, world = world_info(group)
    self.bounds, self.rank, self.world = bounds, rank, world
    self.device = torch.device(device)
    self.local = local_shard
    shards = [dict(row_begin=bounds[r], row_end=bounds[r + 1]) for r in range(world)]
    for key in ('indptr', 'indices', 'eids', 'weights'):
      t = local_shard.get(key)
      present = t is not None
      if not present:
        continue
      peers = exchange_peer_tensors(t, group)
      for r in range(world):
        shards[r][key] = peers[r]
    self._peer_shards = shards
    self.graph = Graph.from_shards(shards, self.device.index)

  @property
  def num_nodes(self):
    return self.bounds[-1]


class PartitionedFeature(object):
  """Row-range partitioned feature table readable from every rank (peer HBM loads)."""

  def __init__(self, local_rows: torch.Tensor, bounds: List[int], device: torch.device, group=None):
    rank, world = world_info(group)
    assert local_rows.shape[0] == bounds[rank + 1] - bounds[rank]
    self.bounds, self.rank, self.world = bounds, rank, world
    self.device = torch.device(device)
    self.local = local_rows.to(self.device).contiguous()
    self.peers = exchange_peer_tensors(self.local, group)
    self.unified = UnifiedTensor(self.device.index, self.local.dtype)
    for p in self.peers:
      self.unified.append_shared_tensor(p)

  @property
  def table(self):
    return self.unified._table()

  def __getitem__(self, ids):
    return self.unified[ids]

  @property
  def shape(self):
    return self.unified.shape
