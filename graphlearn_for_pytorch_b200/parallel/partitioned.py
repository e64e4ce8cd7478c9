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

  def __init__(self, local_shard: dict, bounds: List[int], device: torch.device, group=None,
               replicate_indptr: bool = True, replicate_topology: bool = False):
    """replicate_indptr: keep a local copy of every shard's row-pointer array (8 B per node in
    total) so that degree / row-extent lookups never cross NVLink; only the sampled column
    ids (and edge ids / weights) are read from the owner.
    replicate_topology: pull EVERY shard's arrays over NVLink once at setup and keep them in local
    HBM (the layout the reference uses on a multi-GPU box: `graph_mode='CUDA'` gives each trainer
    process its own copy of the CSR, examples/multi_gpu/train_sage_ogbn_papers100m.py:99-108).  Worth
    it whenever the topology fits next to the features -- 0.5 GB at products shape, 6.5 GB at
    papers100M shape, of 180 GB; sampling then never leaves the GPU.  The shard structure (and with it
    the in-kernel owner lookup) is unchanged, only the pointers are local."""
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
        if r != rank and (replicate_topology or (key == 'indptr' and replicate_indptr)):
          shards[r][key] = peers[r].clone()      # one NVLink pull at setup, local reads afterwards
        else:
          shards[r][key] = peers[r]
    self._peer_shards = shards
    self.replicated = bool(replicate_topology)
    if replicate_topology:
      import torch.distributed as dist
      torch.cuda.synchronize(self.device)
      if dist.is_initialized():
        dist.barrier(group=group)                # nobody frees a shard a peer is still pulling
    self.graph = Graph.from_shards(shards, self.device.index)

  @property
  def num_nodes(self):
    return self.bounds[-1]


class PartitionedFeature(object):
  """Row-range partitioned feature table readable from every rank (peer HBM loads), with an
  optional **replicated hot cache**: when ids are hotness-ordered (data.sort_by_in_degree puts
  the hottest rows first) the first `hot_rows` rows are replicated into every rank's local HBM,
  so the most frequently gathered rows never cross NVLink.

  The replica is filled either through an NVSwitch multicast object (`multimem.st`: the owner
  writes each hot row once and the switch delivers it to all replicas; needs
  torch.distributed._symmetric_memory with multicast support) or, as a fallback, by pulling the
  hot slices from the owners' peer-mapped shards.  This replaces the reference's per-DeviceGroup
  host->device replication of the hot part (data/feature.py:185-199).
  """

  def __init__(self, local_rows: torch.Tensor, bounds: List[int], device: torch.device, group=None,
               hot_rows: int = 0, use_multicast: bool = True, hot_per_rank: int = 0, full_replica: bool = False):
    """full_replica: the whole table fits next to everything else on every GPU (0.6 GB at products shape of 180 GB):
    every rank multicasts its shard once into a symmetric [N, F] buffer in global id order and the table has ONE
    local part -- no owner lookup, no NVLink traffic for features afterwards.
    hot_rows: replicate the global prefix [0, hot_rows) (ids globally hotness-ordered).
    hot_per_rank: replicate the first `hot_per_rank` rows of EVERY rank's range (ids dealt
    round-robin by hotness, see `hotness_balanced_order`): balanced ownership + hot replica."""
    rank, world = world_info(group)
    assert local_rows.shape[0] == bounds[rank + 1] - bounds[rank]
    self.bounds, self.rank, self.world = bounds, rank, world
    self.device = torch.device(device)
    self.local = local_rows.to(self.device).contiguous()
    self.peers = exchange_peer_tensors(self.local, group)
    self.local = self.peers[rank]
    self.hot_rows = int(min(max(hot_rows, 0), bounds[-1])) if world > 1 else 0
    self.hot_per_rank = int(min(hot_per_rank, min(bounds[r + 1] - bounds[r] for r in range(world)))) \
        if world > 1 else 0
    self.replica = None
    self.fill_mode = None
    self.unified = UnifiedTensor(self.device.index, self.local.dtype)
    if full_replica and world > 1:
      self.replica = self._build_full_replica(group, use_multicast)
      self.unified.append_shared_tensor(self.replica)
      self.hot_per_rank, self.hot_rows = 0, bounds[-1]
      return
    if self.hot_per_rank > 0:
      assert 2 * world <= 16, 'per-rank hot replicas need 2 table parts per rank'
      h = self.hot_per_rank
      self.replica = self._build_replica_per_rank(group, use_multicast)
      for r, p in enumerate(self.peers):
        self.unified.append_shared_tensor(self.replica[r * h:(r + 1) * h])   # hot head of rank r: local
        if bounds[r + 1] - bounds[r] > h:
          self.unified.append_shared_tensor(p[h:], remote=(r != rank))       # cold tail: owner's HBM
        else:
          self.unified.append_shared_tensor(p[:0])
      return
    if self.hot_rows > 0:
      self.replica = self._build_replica(group, use_multicast)
      self.unified.append_shared_tensor(self.replica)                  # rows [0, H): local replica
    H = self.hot_rows
    for r, p in enumerate(self.peers):
      lo = max(bounds[r], H)
      if lo >= bounds[r + 1]:
        continue
      self.unified.append_shared_tensor(p[lo - bounds[r]:], remote=(r != rank))   # cold rows stay partitioned

  def _build_replica(self, group, use_multicast: bool) -> torch.Tensor:
    import torch.distributed as dist
    H, F = self.hot_rows, self.local.shape[1:]
    b, e = self.bounds[self.rank], self.bounds[self.rank + 1]
    lo, hi = min(max(b, 0), H), min(e, H)            # hot rows owned by this rank
    if use_multicast:
      try:
        import torch.distributed._symmetric_memory as symm_mem
        rep = symm_mem.empty((H, *F), dtype=self.local.dtype, device=self.device)
        hdl = symm_mem.rendezvous(rep, group=group if group is not None else dist.group.WORLD)
        mc = int(getattr(hdl, 'multicast_ptr', 0) or 0)
        row_bytes = rep[0].numel() * rep.element_size()
        if mc != 0 and row_bytes % 16 == 0:
          if hi > lo:
            require_native().multimem_copy(self.local[lo - b:hi - b].contiguous(), mc, lo * row_bytes)
          torch.cuda.synchronize(self.device)
          dist.barrier(group=group)
          self.fill_mode = 'nvswitch-multicast'
          self._symm_handle = hdl
          return rep
      except Exception as ex:  # noqa: BLE001 - multicast is optional
        self._multicast_error = repr(ex)
    rep = torch.empty((H, *F), dtype=self.local.dtype, device=self.device)
    for r, p in enumerate(self.peers):
      plo, phi = min(self.bounds[r], H), min(self.bounds[r + 1], H)
      if phi > plo:
        rep[plo:phi].copy_(p[plo - self.bounds[r]:phi - self.bounds[r]])   # NVLink pull
    torch.cuda.synchronize(self.device)
    dist.barrier(group=group)
    self.fill_mode = 'peer-pull'
    return rep

  def _build_full_replica(self, group, use_multicast: bool) -> torch.Tensor:
    import torch.distributed as dist
    N, F = self.bounds[-1], self.local.shape[1:]
    b = self.bounds[self.rank]
    if use_multicast:
      try:
        import torch.distributed._symmetric_memory as symm_mem
        rep = symm_mem.empty((N, *F), dtype=self.local.dtype, device=self.device)
        hdl = symm_mem.rendezvous(rep, group=group if group is not None else dist.group.WORLD)
        mc = int(getattr(hdl, 'multicast_ptr', 0) or 0)
        row_bytes = rep[0].numel() * rep.element_size()
        if mc != 0 and row_bytes % 16 == 0:
          # every rank multicasts its own shard once; NVSwitch delivers it to all replicas
          if self.local.shape[0] > 0:
            require_native().multimem_copy(self.local.contiguous(), mc, b * row_bytes)
          torch.cuda.synchronize(self.device)
          dist.barrier(group=group)
          self.fill_mode = 'nvswitch-multicast'
          self._symm_handle = hdl
          return rep
      except Exception as ex:  # noqa: BLE001 - multicast is optional
        self._multicast_error = repr(ex)
    rep = torch.empty((N, *F), dtype=self.local.dtype, device=self.device)
    for r, p in enumerate(self.peers):
      rep[self.bounds[r]:self.bounds[r + 1]].copy_(p)                    # NVLink pull
    torch.cuda.synchronize(self.device)
    dist.barrier(group=group)
    self.fill_mode = 'peer-pull'
    return rep

  def _build_replica_per_rank(self, group, use_multicast: bool) -> torch.Tensor:
    import torch.distributed as dist
    h, W, F = self.hot_per_rank, self.world, self.local.shape[1:]
    if use_multicast:
      try:
        import torch.distributed._symmetric_memory as symm_mem
        rep = symm_mem.empty((W * h, *F), dtype=self.local.dtype, device=self.device)
        hdl = symm_mem.rendezvous(rep, group=group if group is not None else dist.group.WORLD)
        mc = int(getattr(hdl, 'multicast_ptr', 0) or 0)
        row_bytes = rep[0].numel() * rep.element_size()
        if mc != 0 and row_bytes % 16 == 0:
          # every rank multicasts its own hot head once; NVSwitch delivers it to all replicas
          require_native().multimem_copy(self.local[:h].contiguous(), mc, self.rank * h * row_bytes)
          torch.cuda.synchronize(self.device)
          dist.barrier(group=group)
          self.fill_mode = 'nvswitch-multicast'
          self._symm_handle = hdl
          return rep
      except Exception as ex:  # noqa: BLE001
        self._multicast_error = repr(ex)
    rep = torch.empty((W * h, *F), dtype=self.local.dtype, device=self.device)
    for r, p in enumerate(self.peers):
      rep[r * h:(r + 1) * h].copy_(p[:h])
    torch.cuda.synchronize(self.device)
    dist.barrier(group=group)
    self.fill_mode = 'peer-pull'
    return rep

  @property
  def table(self):
    return self.unified._table()

  def __getitem__(self, ids):
    return self.unified[ids]

  @property
  def shape(self):
    return self.unified.shape


def hotness_balanced_order(hotness: torch.Tensor, world: int):
  """Relabelling that makes range partitions balanced *and* hot-first: nodes are sorted by
  hotness (e.g. degree or sample_prob) and dealt round-robin to the ranks, so every rank's id
  range starts with its hottest rows.  -> (old2new [N], bounds [world+1])."""
  n = hotness.numel()
  order = torch.argsort(hotness, descending=True, stable=True)
  counts = [(n - r + world - 1) // world for r in range(world)]
  bounds = [0]
  for c in counts:
    bounds.append(bounds[-1] + c)
  i = torch.arange(n, device=hotness.device)
  b = torch.tensor(bounds[:-1], device=hotness.device, dtype=torch.int64)
  new_pos = b[i % world] + i // world
  old2new = torch.empty(n, dtype=torch.int64, device=hotness.device)
  old2new[order] = new_pos
  return old2new, bounds


def partition_hetero_graph(topo_dict: dict, num_nodes: dict, rank: int, world: int, device: torch.device,
                           edge_dir: str = 'out', group=None):
  """Range-partition every relation of a heterogeneous graph by the node type its CSR/CSC rows
  belong to and map all shards on every rank.  -> ({etype: Graph}, {ntype: bounds})."""
  bounds = {nt: range_bounds(n, world) for nt, n in num_nodes.items()}
  graphs, keep = {}, []
  for et, topo in topo_dict.items():
    row_type = et[0] if edge_dir == 'out' else et[2]
    shard = shard_topology(topo, bounds[row_type], rank, device)
    pg = PartitionedGraph(shard, bounds[row_type], device, group)
    keep.append(pg)
    graphs[et] = pg.graph
  return graphs, bounds, keep
