"""DistNeighborSampler: multi-hop sampling over a partitioned graph.

Parity: reference python/distributed/dist_neighbor_sampler.py:96-807 (async hop loop,
partition scatter / remote one-hop / stitch, link + subgraph sampling, message collation
with labels and features).  Data planes:
  * p2p : the dataset's Graph already spans every rank's shard over NVLink; the whole
          multi-hop sample is the single-process arena sampler (no RPC, no stitch);
  * rpc : seeds are split by the node partition book, remote partitions are sampled by
          their owners through RPC, partial results are stitched back in seed order.
"""
import asyncio
import math
from typing import Dict, List, Optional, Tuple, Union

import torch

from ..channel import ChannelBase, SampleMessage
from ..ops import require_native
from ..ops.tables import IdTable
from ..sampler import (EdgeSamplerInput, HeteroSamplerOutput, NegativeSampling, NeighborOutput, NeighborSampler,
                       NodeSamplerInput, SamplerOutput, SamplingConfig, SamplingType)
from ..typing import EdgeType, NodeType, as_str, reverse_edge_type
from ..utils.common import count_dict
from .dist_dataset import DistDataset
from .dist_feature import DistFeature
from .dist_graph import DistGraph
from .event_loop import ConcurrentEventLoop, wrap_torch_future
from .rpc import (RpcCalleeBase, RpcDataPartitionRouter, rpc_is_initialized, rpc_register, rpc_request_async,
                  rpc_sync_data_partitions)


class RpcSamplingCallee(RpcCalleeBase):
  """Serves one-hop sampling requests for the ids this partition owns."""

  def __init__(self, sampler: NeighborSampler, device: torch.device):
    self.sampler, self.device = sampler, device

  def call(self, ids: torch.Tensor, req_num: int, etype=None, stream: int = 0):
    out = self.sampler.sample_one_hop(ids.to(self.device), req_num, etype, stream=stream)
    return out.nbr.cpu(), out.nbr_num.cpu(), (out.edge.cpu() if out.edge is not None else None)


class RpcSubGraphCallee(RpcCalleeBase):
  """Returns the induced edges among `nodes` that live in this partition (global ids)."""

  def __init__(self, sampler: NeighborSampler, device: torch.device):
    self.sampler, self.device = sampler, device

  def call(self, nodes: torch.Tensor, with_edge: bool):
    node, rows, cols, eids, _ = self.sampler.node_subgraph(nodes.to(self.device))
    return node[rows].cpu(), node[cols].cpu(), (eids.cpu() if eids is not None else None)


def stitch_one_hop(n_seeds: int, parts: List[Tuple[torch.Tensor, NeighborOutput]], device, with_edge: bool):
  """Merge per-partition one-hop results back into seed order (torch ops on `device`;
  the native CPU variant is csrc/cpu/cpu_ops.cc::cpu_stitch)."""
  if device.type == 'cpu':
    nat = require_native()
    idx = [p[0].cpu().contiguous() for p in parts]
    nbr, num, eid = nat.cpu_stitch(n_seeds, idx, [p[1].nbr.cpu().contiguous() for p in parts],
                                   [p[1].nbr_num.cpu().contiguous() for p in parts],
                                   [p[1].edge.cpu().contiguous() for p in parts] if with_edge else [])
    return NeighborOutput(nbr, num, eid if with_edge else None)
  counts = torch.zeros(n_seeds, dtype=torch.int64, device=device)
  for idx, out in parts:
    counts[idx.to(device)] = out.nbr_num.to(device)
  offs = torch.cumsum(counts, 0) - counts
  total = int(counts.sum())
  nbr = torch.empty(total, dtype=torch.int64, device=device)
  eid = torch.empty(total, dtype=torch.int64, device=device) if with_edge else None
  for idx, out in parts:
    num = out.nbr_num.to(device)
    if num.numel() == 0 or int(num.sum()) == 0:
      continue
    start = torch.repeat_interleave(offs[idx.to(device)], num)
    within = torch.arange(int(num.sum()), device=device) - torch.repeat_interleave(torch.cumsum(num, 0) - num, num)
    nbr[start + within] = out.nbr.to(device)
    if with_edge:
      eid[start + within] = out.edge.to(device)
  return NeighborOutput(nbr, counts, eid)


class DistNeighborSampler(ConcurrentEventLoop):
  """Args follow the reference: data, num_neighbors, with_edge, with_neg, with_weight,
  edge_dir, collect_features, channel (None = return results to the caller), use_all2all,
  concurrency, device, seed."""

  def __init__(self, data: DistDataset, num_neighbors=None, with_edge: bool = False, with_neg: bool = False,
               with_weight: bool = False, edge_dir: str = 'out', collect_features: bool = False,
               channel: Optional[ChannelBase] = None, use_all2all: bool = False, concurrency: int = 1,
               device: Optional[torch.device] = None, seed: Optional[int] = None):
    super().__init__(concurrency)
    self.data = data
    self.num_neighbors = num_neighbors
    self.with_edge, self.with_neg, self.with_weight = with_edge, with_neg, with_weight
    self.edge_dir = edge_dir
    self.collect_features = collect_features
    self.channel = channel
    self.use_all2all = use_all2all
    self.data_plane = getattr(data, 'data_plane', 'rpc')
    self.sampler = NeighborSampler(data.graph, num_neighbors, device, with_edge, with_neg, with_weight,
                                   edge_dir=edge_dir, seed=seed)
    self.device = self.sampler.device
    self.data_cls = 'hetero' if isinstance(data.graph, dict) else 'homo'
    self.num_partitions, self.partition_idx = data.num_partitions, data.partition_idx
    self.dist_graph = DistGraph(self.num_partitions, self.partition_idx, data.graph, data.node_pb, data.edge_pb)
    self.rpc_router = None
    self.rpc_sample_callee_id = self.rpc_subgraph_callee_id = None
    self.dist_node_feature = self.dist_edge_feature = None
    remote = self.data_plane == 'rpc' and self.num_partitions > 1
    if remote:
      assert rpc_is_initialized(), 'rpc plane with several partitions needs init_rpc()'
      self.rpc_router = RpcDataPartitionRouter(rpc_sync_data_partitions(self.num_partitions, self.partition_idx))
    if data.node_features is not None and self.data_plane == 'rpc':
      self.dist_node_feature = DistFeature(self.num_partitions, self.partition_idx, data.node_features,
                                           data.node_feat_pb, local_only=not remote,
                                           rpc_router=self.rpc_router, device=self.device)
    if data.edge_features is not None and self.data_plane == 'rpc':
      self.dist_edge_feature = DistFeature(self.num_partitions, self.partition_idx, data.edge_features,
                                           data.edge_feat_pb, local_only=not remote,
                                           rpc_router=self.rpc_router, device=self.device)
    if remote:
      self.rpc_sample_callee_id = rpc_register(RpcSamplingCallee(self.sampler, self.device))
      self.rpc_subgraph_callee_id = rpc_register(RpcSubGraphCallee(self.sampler, self.device))
    self._remote = remote

  # ------------------------------------------------------------------ public entry points
  def sample_from_nodes(self, inputs: NodeSamplerInput) -> Optional[SampleMessage]:
    inputs = NodeSamplerInput.cast(inputs)
    return self._dispatch(self._send_adapter(self._sample_from_nodes, inputs))

  def sample_from_edges(self, inputs: EdgeSamplerInput) -> Optional[SampleMessage]:
    inputs = EdgeSamplerInput.cast(inputs)
    return self._dispatch(self._send_adapter(self._sample_from_edges, inputs))

  def subgraph(self, inputs: NodeSamplerInput) -> Optional[SampleMessage]:
    inputs = NodeSamplerInput.cast(inputs)
    return self._dispatch(self._send_adapter(self._subgraph, inputs))

  def _dispatch(self, coro):
    if self.channel is None:
      return self.run_task(coro)
    self.add_task(coro)
    return None

  async def _send_adapter(self, fn, *args, **kwargs):
    out = await fn(*args, **kwargs)
    msg = await self._colloate_fn(out)
    if self.channel is None:
      return msg
    self.channel.send(msg)
    return None

  # ------------------------------------------------------------------ one hop
  async def _sample_one_hop(self, srcs: torch.Tensor, num_nbr: int, etype: Optional[EdgeType], stream: int
                            ) -> NeighborOutput:
    if not self._remote:
      return self.sampler.sample_one_hop(srcs, num_nbr, etype, stream=stream)
    src_t = None
    if etype is not None:
      src_t = etype[0] if self.edge_dir == 'out' else etype[-1]
    owners = self.dist_graph.get_node_partitions(srcs, src_t).to(srcs.device)
    parts, futs = [], []
    order = [self.partition_idx] + [p for p in range(self.num_partitions) if p != self.partition_idx]
    for p in order:
      idx = torch.nonzero(owners == p, as_tuple=False).view(-1)
      if idx.numel() == 0:
        continue
      ids = srcs[idx]
      if p == self.partition_idx:
        parts.append((idx, self.sampler.sample_one_hop(ids, num_nbr, etype, stream=stream)))
      else:
        to = self.rpc_router.get_to_worker(p)
        f = rpc_request_async(to, self.rpc_sample_callee_id, args=(ids.cpu(), num_nbr, etype, stream))
        futs.append((idx, wrap_torch_future(f)))
    for idx, f in futs:
      nbr, num, eid = await f
      parts.append((idx, NeighborOutput(nbr, num, eid)))
    return stitch_one_hop(srcs.numel(), parts, self.device, self.with_edge)

  # ------------------------------------------------------------------ from nodes
  async def _sample_from_nodes(self, inputs: NodeSamplerInput):
    seeds = inputs.node.to(self.device, dtype=torch.int64)
    if self.data_cls == 'hetero':
      out = await self._hetero_from_seeds({inputs.input_type: seeds})
      out.input_type = inputs.input_type
      return out
    return await self._homo_from_seeds(seeds)

  async def _homo_from_seeds(self, seeds: torch.Tensor) -> SamplerOutput:
    if not self._remote:
      return self.sampler._sample_from_nodes(seeds)
    s = self.sampler
    table = IdTable(self.device, s._max_sampled_nodes(seeds.numel()))
    table.init(seeds)
    n0 = table.size()
    frontier = table.keys(0)
    frontier_local = torch.arange(n0, dtype=torch.int64, device=self.device)
    rows, cols, eids, nn, ne = [], [], [], [n0], []
    stream = s._next_stream(len(self.num_neighbors))
    for h, k in enumerate(self.num_neighbors):
      out = await self._sample_one_hop(frontier, k, None, stream + h)
      if out.nbr.numel() == 0:
        break
      before = table.size()
      nbr_local = table.insert(out.nbr)
      rows.append(nbr_local)
      cols.append(torch.repeat_interleave(frontier_local, out.nbr_num.to(self.device)))
      if out.edge is not None:
        eids.append(out.edge)
      after = table.size()
      nn.append(after - before)
      ne.append(int(out.nbr.numel()))
      frontier = table.keys(before)
      frontier_local = torch.arange(before, after, dtype=torch.int64, device=self.device)
      if frontier.numel() == 0:
        break
    e = torch.empty(0, dtype=torch.int64, device=self.device)
    node = table.keys(0)
    return SamplerOutput(node=node, row=torch.cat(rows) if rows else e, col=torch.cat(cols) if cols else e.clone(),
                         edge=(torch.cat(eids) if eids else e.clone()) if self.with_edge else None,
                         batch=node[:n0], num_sampled_nodes=nn, num_sampled_edges=ne, device=self.device)

  async def _hetero_from_seeds(self, seeds_dict: Dict[NodeType, torch.Tensor]) -> HeteroSamplerOutput:
    if not self._remote:
      return self.sampler._hetero_sample_from_nodes(seeds_dict)
    s = self.sampler
    cap = s._hetero_table_cap(sum(v.numel() for v in seeds_dict.values()))
    tables: Dict[NodeType, IdTable] = {}

    n_seed_total = sum(v.numel() for v in seeds_dict.values())

    def table_of(nt):
      if nt not in tables:
        # global node count of the type, when the partition book can tell (the local shard cannot)
        pb = self.data.node_pb.get(nt) if isinstance(self.data.node_pb, dict) else None
        total = None
        if isinstance(pb, torch.Tensor):
          total = pb.numel()
        elif hasattr(pb, 'partition_bounds'):
          total = int(pb.partition_bounds[-1])
        elif hasattr(pb, '_ranges'):
          total = int(pb._ranges[-1])
        bound = s._hetero_type_bound(nt, n_seed_total, total) if total is not None else None
        tables[nt] = IdTable(self.device, min(cap, bound) if bound else cap)
      return tables[nt]
    src_dict, src_local, num_nodes, num_edges = {}, {}, {}, {}
    for nt, sd in seeds_dict.items():
      t = table_of(nt)
      t.init(sd)
      src_dict[nt] = t.keys(0)
      src_local[nt] = torch.arange(t.size(), dtype=torch.int64, device=self.device)
    count_dict(src_dict, num_nodes, 1)
    batch = {nt: v.clone() for nt, v in src_dict.items()}
    rows, cols, eids = {}, {}, {}
    n_et = max(1, len(s.edge_types))
    stream = s._next_stream(s.num_hops * n_et)
    for h in range(s.num_hops):
      before = {nt: t.size() for nt, t in tables.items()}
      tasks = []
      for ei, etype in enumerate(s.edge_types):
        src_t, nbr_t = s._etype_ends(etype)
        src = src_dict.get(src_t)
        if src is None or src.numel() == 0:
          continue
        tasks.append((etype, src_t, nbr_t,
                      asyncio.ensure_future(self._sample_one_hop(src, s.num_neighbors[etype][h], etype,
                                                                 stream + h * n_et + ei))))
      hop_edges = {}
      for etype, src_t, nbr_t, task in tasks:
        out = await task
        if out.nbr.numel() == 0:
          continue
        nbr_local = table_of(nbr_t).insert(out.nbr)
        s_local = torch.repeat_interleave(src_local[src_t], out.nbr_num.to(self.device))
        key = reverse_edge_type(etype) if self.edge_dir == 'out' else etype
        rows.setdefault(key, []).append(nbr_local)
        cols.setdefault(key, []).append(s_local)
        if out.edge is not None:
          eids.setdefault(key, []).append(out.edge)
        hop_edges[key] = out.nbr
      new_src, new_local = {}, {}
      for nt, t in tables.items():
        b = before.get(nt, 0)
        if t.size() > b:
          new_src[nt] = t.keys(b)
          new_local[nt] = torch.arange(b, t.size(), dtype=torch.int64, device=self.device)
      count_dict(new_src, num_nodes, h + 2)
      count_dict(hop_edges, num_edges, h + 1)
      src_dict, src_local = new_src, new_local
      if not new_src:
        break
    out_types = [reverse_edge_type(et) if self.edge_dir == 'out' else et for et in s.edge_types]
    return HeteroSamplerOutput(
      node={nt: t.keys(0) for nt, t in tables.items()},
      row={k: torch.cat(v) for k, v in rows.items()}, col={k: torch.cat(v) for k, v in cols.items()},
      edge={k: torch.cat(v) for k, v in eids.items()} if self.with_edge else None, batch=batch,
      num_sampled_nodes={k: torch.as_tensor(v, dtype=torch.int64) for k, v in num_nodes.items()},
      num_sampled_edges={k: torch.as_tensor(v, dtype=torch.int64) for k, v in num_edges.items()},
      edge_types=out_types, device=self.device)

  # ------------------------------------------------------------------ from edges
  async def _sample_from_edges(self, inputs: EdgeSamplerInput):
    """Negatives are drawn against the *local* graph view: strict on the p2p plane (the
    view spans every shard), local-partition-only on the rpc plane like the reference
    (dist_neighbor_sampler.py:411-413)."""
    src = inputs.row.to(self.device, dtype=torch.int64)
    dst = inputs.col.to(self.device, dtype=torch.int64)
    edge_label = inputs.label.to(self.device) if inputs.label is not None else None
    input_type, neg = inputs.input_type, inputs.neg_sampling
    num_pos = src.numel()
    if neg is not None:
      self.sampler.with_neg = True
      self.sampler.lazy_init_sampler()
      self.sampler.lazy_init_neg_sampler()
      ns = self.sampler._neg_sampler if self.data_cls == 'homo' else self.sampler._neg_sampler[input_type]
      num_neg = math.ceil(num_pos * neg.amount)
      pair = ns.sample(num_neg, padding=True).to(self.device)
      if neg.is_binary():
        src, dst = torch.cat([src, pair[0]]), torch.cat([dst, pair[1]])
        if edge_label is None:
          edge_label = torch.ones(num_pos, device=self.device)
        edge_label = torch.cat([edge_label, edge_label.new_zeros((pair.shape[1],) + tuple(edge_label.shape[1:]))])
      else:
        dst = torch.cat([dst, pair[1]])
    if self.data_cls == 'homo':
      seed, inverse = torch.unique(torch.cat([src, dst]), return_inverse=True)
      out = await self._homo_from_seeds(seed)
      NeighborSampler._attach_link_metadata(out, neg, inverse, num_pos, edge_label, src.numel())
      return out
    src_t, dst_t = input_type[0], input_type[-1]
    if src_t == dst_t:
      seed, inverse = torch.unique(torch.cat([src, dst]), return_inverse=True)
      seeds_dict = {src_t: seed}
      inv_src, inv_dst = inverse[:src.numel()], inverse[src.numel():]
    else:
      s_seed, inv_src = torch.unique(src, return_inverse=True)
      d_seed, inv_dst = torch.unique(dst, return_inverse=True)
      seeds_dict = {src_t: s_seed, dst_t: d_seed}
    out = await self._hetero_from_seeds(seeds_dict)
    out.input_type = input_type
    if neg is None or neg.is_binary():
      out.metadata = {'edge_label_index': torch.stack([inv_src, inv_dst]), 'edge_label': edge_label}
    else:
      out.metadata = {'src_index': inv_src[:num_pos], 'dst_pos_index': inv_dst[:num_pos],
                      'dst_neg_index': inv_dst[num_pos:].view(num_pos, -1)}
    return out

  # ------------------------------------------------------------------ subgraph
  async def _subgraph(self, inputs: NodeSamplerInput) -> SamplerOutput:
    if self.data_cls != 'homo':
      raise NotImplementedError('distributed subgraph sampling supports homogeneous graphs')
    seeds = inputs.node.to(self.device, dtype=torch.int64)
    if not self._remote:
      out = self.sampler.subgraph(NodeSamplerInput(seeds))
      out.batch = seeds
      return out
    nodes = [seeds]
    if self.num_neighbors is not None:
      frontier = torch.unique(seeds)
      stream = self.sampler._next_stream(len(self.num_neighbors))
      for h, k in enumerate(self.num_neighbors):
        nbr = (await self._sample_one_hop(frontier, k, None, stream + h)).nbr
        if nbr.numel() == 0:
          break
        frontier = torch.unique(nbr)
        nodes.append(frontier)
    node = torch.unique(torch.cat(nodes))          # ascending ids, like the single-machine subgraph()
    owners = self.dist_graph.get_node_partitions(node, None)
    futs, rows_g, cols_g, eids_g = [], [], [], []
    for p in range(self.num_partitions):
      if not bool((owners == p).any()):
        continue
      if p == self.partition_idx:
        n2, r2, c2, e2, _ = self.sampler.node_subgraph(node)
        rows_g.append(n2[r2]); cols_g.append(n2[c2])
        if e2 is not None:
          eids_g.append(e2)
      else:
        to = self.rpc_router.get_to_worker(p)
        futs.append(wrap_torch_future(rpc_request_async(to, self.rpc_subgraph_callee_id,
                                                        args=(node.cpu(), self.with_edge))))
    for f in futs:
      r, c, e = await f
      rows_g.append(r.to(self.device)); cols_g.append(c.to(self.device))
      if e is not None:
        eids_g.append(e.to(self.device))
    empty = torch.empty(0, dtype=torch.int64, device=self.device)
    rg = torch.cat(rows_g) if rows_g else empty
    cg = torch.cat(cols_g) if cols_g else empty
    return SamplerOutput(node=node, row=torch.searchsorted(node, rg), col=torch.searchsorted(node, cg),
                         edge=(torch.cat(eids_g) if eids_g else empty) if self.with_edge else None,
                         batch=seeds, device=self.device, metadata=torch.searchsorted(node, seeds))

  # ------------------------------------------------------------------ message collation
  async def _get_node_feats(self, ids: torch.Tensor, ntype=None):
    if self.data_plane == 'p2p':
      feat = self.data.node_features
      feat = feat[ntype] if isinstance(feat, dict) else feat
      return feat[ids] if feat is not None else None
    if self.dist_node_feature is None:
      return None
    if self.use_all2all:
      return self.dist_node_feature.get_all2all(ids, ntype)
    return await wrap_torch_future(self.dist_node_feature.async_get(ids, ntype))

  async def _get_edge_feats(self, eids: torch.Tensor, etype=None):
    if self.dist_edge_feature is None:
      return None
    return await wrap_torch_future(self.dist_edge_feature.async_get(eids, etype))

  async def _colloate_fn(self, output: Union[SamplerOutput, HeteroSamplerOutput]) -> SampleMessage:
    """-> flat Dict[str, Tensor] (wire keys: SURVEY.md Appendix B)."""
    msg: SampleMessage = {}
    is_hetero = isinstance(output, HeteroSamplerOutput)
    msg['#IS_HETERO'] = torch.tensor([int(is_hetero)])
    md = output.metadata
    if isinstance(md, dict):
      for k, v in md.items():
        if v is not None:
          msg[f'#META.{k}'] = v
    elif md is not None:
      msg['#META.mapping'] = md
    labels = self.data.node_labels
    if is_hetero:
      for nt, ids in output.node.items():
        msg[f'{as_str(nt)}.ids'] = ids
        msg[f'{as_str(nt)}.num_sampled_nodes'] = torch.as_tensor(output.num_sampled_nodes.get(nt, []))
        if output.batch is not None and nt in output.batch:
          msg[f'{as_str(nt)}.batch'] = output.batch[nt]
        lab = labels.get(nt) if isinstance(labels, dict) else None
        if lab is not None:
          msg[f'{as_str(nt)}.nlabels'] = lab[ids.to(lab.device)]
        if self.collect_features:
          x = await self._get_node_feats(ids, nt)
          if x is not None:
            msg[f'{as_str(nt)}.nfeats'] = x
      for et, rows in output.row.items():
        k = as_str(et)
        msg[f'{k}.rows'], msg[f'{k}.cols'] = rows, output.col[et]
        msg[f'{k}.num_sampled_edges'] = torch.as_tensor(output.num_sampled_edges.get(et, []))
        if output.edge is not None and et in output.edge:
          msg[f'{k}.eids'] = output.edge[et]
          if self.collect_features:
            orig = reverse_edge_type(et) if self.edge_dir == 'out' else et
            ef = await self._get_edge_feats(output.edge[et], orig)
            if ef is not None:
              msg[f'{k}.efeats'] = ef
      if output.input_type is not None:
        t = output.input_type
        msg['#META.input_type'] = torch.tensor(list(as_str(t).encode()), dtype=torch.uint8)
      return msg
    msg['ids'], msg['rows'], msg['cols'] = output.node, output.row, output.col
    if output.num_sampled_nodes is not None:
      msg['num_sampled_nodes'] = torch.as_tensor(output.num_sampled_nodes)
      msg['num_sampled_edges'] = torch.as_tensor(output.num_sampled_edges)
    if output.batch is not None:
      msg['batch'] = output.batch
    if output.edge is not None:
      msg['eids'] = output.edge
    if isinstance(labels, torch.Tensor):
      msg['nlabels'] = labels[output.node.to(labels.device)]
    if self.collect_features:
      x = await self._get_node_feats(output.node)
      if x is not None:
        msg['nfeats'] = x
      if output.edge is not None:
        ef = await self._get_edge_feats(output.edge)
        if ef is not None:
          msg['efeats'] = ef
    return msg
