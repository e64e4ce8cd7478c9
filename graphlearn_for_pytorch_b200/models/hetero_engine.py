"""`HeteroSageEngine`: device-resident relational GraphSAGE (R-SAGE) training step for heterogeneous graphs.

The reference trains IGBH with a PyG `HeteroConv({etype: SAGEConv})` fed by its hetero `NeighborLoader`
(examples/igbh/rgnn.py:22-81, examples/igbh/train_rgnn_multi_gpu.py:80-142): per hop and per edge type a sampling
call, a D2H size read, an inducer call, then per-relation scatter/linear autograd ops.  Here the whole step is a
fixed launch sequence over static buffers, captured into ONE CUDA graph:

  sample   native `HeteroArena`: one grouped launch per hop over all relations, per-type id tables, no host sync
  layer l  for every destination type t that can hold targets at this depth:
             A_t = [ mean_{r -> t} x_src(r)  ...  | x_t ]   one ELL aggregation kernel per incoming relation,
                                                           writing its column block (layer 1 reads the feature
                                                           tables -- local or peer HBM -- in place)
             Z_t = act(A_t . W_t^T + b_t)                  TMA-fed tcgen05 GEMM, rows from the device counter
  loss     fused log-softmax / NLL on the seed type (labels looked up through the batch's node list)
  backward per (layer, type): dW (split-K, fp32 red-add) + dA in one tcgen05 launch, then per relation the ELL
           scatter of its dA column block into the source type's fp32 gradient rows, ReLU mask + bf16 cast
  update   fused flat Adam (fp32 master weights, bf16 shadow copy)

`W_t = [W_r1 | W_r2 | ... | W_self]` concatenates the relation weights of a destination type, so summing the
per-relation SAGEConv outputs (HeteroConv aggr='sum') is a single GEMM.  Activation is ReLU (the reference example
uses leaky_relu + dropout; see docs).  Node types are integer coded in sorted order, relations in the order of
`graphs`.
"""
import math
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

from ..ops import require_native

EdgeType = Tuple[str, str, str]


def _round_up(x, m):
  return (x + m - 1) // m * m


class HeteroSageEngine(object):
  """Args:
    graphs: {(src, rel, dst): data.Graph} device graphs (single- or multi-shard), CSC by dst for edge_dir='in',
      CSR by src for edge_dir='out'.
    feature_tables: {node type: native RowTableHandle of bf16 rows indexed by global id} (`UnifiedTensor._table()`,
      `PartitionedFeature.table`).
    labels: int64 [num_nodes[seed_type]].
    in_dim: feature width (int or {type: int}; multiples of 8).
    num_nodes: {type: node count}.
    fanouts: [k per hop] (all relations) or {etype: [k per hop]}.
  """

  def __init__(self, graphs: Dict[EdgeType, object], feature_tables: Dict[str, object], labels: torch.Tensor,
               in_dim: Union[int, Dict[str, int]], num_nodes: Dict[str, int], seed_type: str,
               fanouts: Union[Sequence[int], Dict[EdgeType, Sequence[int]]] = (15, 10, 5), batch_size: int = 1024,
               hidden: int = 512, num_classes: int = 19, edge_dir: str = 'in', lr: float = 1e-3,
               weight_decay: float = 0.0, seed: int = 0, device: Optional[torch.device] = None,
               use_cuda_graph: bool = True, cap_limit: int = 1 << 21, group=None):
    self.nat = require_native()
    self.edge_types = [tuple(e) for e in graphs]
    self.graphs = graphs
    self.edge_dir = edge_dir
    first = next(iter(graphs.values()))
    first.lazy_init()
    self.device = torch.device(device) if device is not None else torch.device('cuda', first.device)
    dev = self.device
    self.seed_type = seed_type
    self.bs = int(batch_size)
    self.hidden, self.C = int(hidden), int(num_classes)
    self.n_out_pad = _round_up(self.C, 64)
    self.lr, self.wd, self.seed = lr, weight_decay, int(seed)
    self.use_cuda_graph = use_cuda_graph
    import torch.distributed as dist
    self.group = group
    self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    self.peer_group = None
    # ---- integer-coded schema -------------------------------------------------------------------------
    ends = [(et[0], et[2]) if edge_dir == 'out' else (et[2], et[0]) for et in self.edge_types]   # (key, nbr)
    self.ntypes = sorted({t for e in ends for t in e} | {seed_type})
    self.tid = {t: i for i, t in enumerate(self.ntypes)}
    self.kt = [self.tid[k] for k, _ in ends]          # type a relation is sampled FROM (its messages' destination)
    self.nt = [self.tid[n] for _, n in ends]          # neighbour type (message source)
    if isinstance(fanouts, dict):
      self.fan = [[int(k) for k in fanouts[et]] for et in self.edge_types]
    else:
      self.fan = [[int(k) for k in fanouts] for _ in self.edge_types]
    self.L = len(self.fan[0])
    assert 1 <= self.L <= 4 and all(len(f) == self.L for f in self.fan)
    NT, NR = len(self.ntypes), len(self.edge_types)
    self.dim_in = {t: int(in_dim[t] if isinstance(in_dim, dict) else in_dim) for t in self.ntypes}
    assert all(d % 8 == 0 for d in self.dim_in.values()) and self.hidden % 64 == 0
    self.feat = feature_tables
    self.labels = labels.to(dev)
    # types that can appear at hop h (static reachability over the schema)
    hop_types = [{self.tid[seed_type]}]
    for h in range(self.L):
      hop_types.append({self.nt[r] for r in range(NR) if self.kt[r] in hop_types[h] and self.fan[r][h] > 0})
    self.hop_types = hop_types
    # layer l (1-based) has targets of the types reachable within hops 0..L-l
    self.targets = [None] + [sorted(set().union(*hop_types[:self.L - l + 1])) for l in range(1, self.L + 1)]
    with torch.cuda.device(dev):
      handles = []
      for et in self.edge_types:
        graphs[et].lazy_init()
        handles.append(graphs[et].graph_handler)
      nn_list = [int(num_nodes[t]) for t in self.ntypes]
      max_seeds = [self.bs if t == seed_type else 0 for t in self.ntypes]
      self.arena = self.nat.HeteroArena(dev.index, NT, handles, self.kt, self.nt, self.fan, nn_list, max_seeds,
                                        False, self.seed, False, False, int(cap_limit), [])
      self.cap_rows = [list(v) for v in self.arena.cap_rows]
      self.cap_nodes = list(self.arena.cap_nodes)
      self._ctr = [self.arena.counters[t * 8:t * 8 + 8] for t in range(NT)]       # per-type hop counters (views)
      self._ovf_idx = self.arena.overflow_index()
      # relations feeding type t at layer l: every relation sampled from t with a positive fan-out in a used hop
      self.in_rel = [None]
      for l in range(1, self.L + 1):
        nh = self.L - l + 1
        # (a relation only counts for the hops at which its key type can actually be in the frontier)
        self.in_rel.append({t: [r for r in range(NR) if self.kt[r] == t and
                                any(self.fan[r][h] > 0 and t in hop_types[h] for h in range(nh))]
                            for t in self.targets[l]})
      self._alloc()
      self._init_params()
      self._seeds = torch.zeros(self.bs, dtype=torch.int64, device=dev)
      self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
      self.correct = torch.zeros(1, dtype=torch.int32, device=dev)
      self.step_dev = torch.zeros(2, dtype=torch.int32, device=dev)
    self._plans = {}
    self._graph = None
    self.step_idx = 0
    self.kernels_per_step = 0

  # ------------------------------------------------------------------ buffers / parameters
  def _cap_T(self, l: int, t: int) -> int:
    return int(min(self.cap_nodes[t], sum(self.cap_rows[t][:self.L - l + 1])))

  def _width_in(self, l: int, t: int) -> int:
    return self.dim_in[self.ntypes[t]] if l == 1 else self.hidden

  def _alloc(self):
    dev, bf, f32 = self.device, torch.bfloat16, torch.float32
    self.A, self.Z, self.dPre, self.dA, self.dH, self.col = [None], [None], [None], [None], [None], [None]
    for l in range(1, self.L + 1):
      n_out = self.hidden if l < self.L else self.n_out_pad
      A, Z, dP, dA, dH, col = {}, {}, {}, {}, {}, {}
      for t in self.targets[l]:
        cap = max(self._cap_T(l, t), 128)
        offs, c = [], 0
        for r in self.in_rel[l][t]:
          offs.append(c)
          c += self._width_in(l, self.nt[r])
        col[t] = (offs, c)                       # mean-block offsets, self-block offset
        K = c + self._width_in(l, t)
        A[t] = torch.zeros(cap, K, dtype=bf, device=dev)
        Z[t] = torch.zeros(cap, n_out, dtype=bf, device=dev)
        dP[t] = torch.zeros(cap, n_out, dtype=bf, device=dev)
        if l > 1:
          dA[t] = torch.zeros(cap, K, dtype=bf, device=dev)
        if l < self.L:
          dH[t] = torch.zeros(cap, n_out, dtype=f32, device=dev)   # gradient wrt Z[l][t], filled by layer l+1
      self.A.append(A); self.Z.append(Z); self.dPre.append(dP); self.dA.append(dA); self.dH.append(dH)
      self.col.append(col)

  def _init_params(self):
    dev = self.device
    self._w_off, self._b_off, sizes = {}, {}, 0
    for l in range(1, self.L + 1):
      for t in self.targets[l]:
        n, k = self.Z[l][t].shape[1], self.A[l][t].shape[1]
        self._w_off[(l, t)] = (sizes, n, k); sizes += n * k
        self._b_off[(l, t)] = (sizes, n); sizes += n
    self.p32 = torch.zeros(sizes, dtype=torch.float32, device=dev)
    g = torch.Generator(device='cpu')
    g.manual_seed(self.seed)
    for (l, t), (off, n, k) in self._w_off.items():
      w = (torch.rand(n, k, generator=g) * 2 - 1) / math.sqrt(max(k // (len(self.in_rel[l][t]) + 1), 1))
      if l == self.L and self.n_out_pad > self.C:
        w[self.C:] = 0
      self.p32[off:off + n * k].copy_(w.flatten())
    pad = (-sizes) % 4
    self.p16 = self.p32.to(torch.bfloat16)
    self.g32 = torch.zeros_like(self.p32)
    if self.world > 1:
      # data parallel over the GPUs of the box: the gradient all-reduce is a peer-HBM read fused into the Adam
      # kernel (csrc/cuda/peer.cu), one NVLink barrier per step, no NCCL call inside the captured step
      import torch.distributed as dist
      from ..parallel.peer import exchange_peer_tensors
      assert pad == 0, 'flat parameter buffer must be a multiple of 4 floats'
      rank = dist.get_rank(self.group)
      peers_g = exchange_peer_tensors(torch.zeros_like(self.p32), self.group)
      flags = exchange_peer_tensors(torch.zeros(2 * self.world, dtype=torch.int32, device=dev), self.group)
      self.g32 = peers_g[rank]
      self.peer_group = self.nat.PeerGroup(dev.index, rank, peers_g, flags)
    self.m = torch.zeros_like(self.p32)
    self.v = torch.zeros_like(self.p32)

  def W(self, l, t):
    off, n, k = self._w_off[(l, t)]
    return self.p16[off:off + n * k].view(n, k)

  def b(self, l, t):
    off, n = self._b_off[(l, t)]
    return self.p16[off:off + n]

  def state_dict(self):
    return {'p32': self.p32.clone(), 'm': self.m.clone(), 'v': self.v.clone(), 'step': int(self.step_dev[0].item()),
            'sample_step': int(self.arena.step.item()), 'step_idx': self.step_idx, 'seed': self.seed}

  def load_state_dict(self, s):
    self.p32.copy_(s['p32']); self.m.copy_(s['m']); self.v.copy_(s['v'])
    self.p16.copy_(self.p32)
    self.step_dev.zero_()
    self.step_dev[0] = int(s['step'])
    self.arena.step.fill_(int(s['sample_step']))
    self.step_idx, self.seed = s['step_idx'], s['seed']

  # ------------------------------------------------------------------ step pieces
  _tally = 0

  def _k(self, n=1):
    self._tally += n

  def _rel_ell(self, r: int, nh: int):
    return [self.arena.ell_of(r, h) for h in range(nh)], [max(self.fan[r][h], 1) for h in range(nh)]

  def _sample(self):
    self.arena.sample([self.tid[self.seed_type]], [self._seeds], self.L * len(self.edge_types))
    self._k(3 + 3 * self.L)

  def _plan(self, kind: str, l: int, t: int):
    key = (kind, l, t)
    pl = self._plans.get(key)
    if pl is None:
      nh = self.L - l + 1
      pl = self.nat.TcGemm(self.device.index)
      if kind == 'fwd':
        pl.add_forward(self.A[l][t], self.W(l, t), self.b(l, t), l < self.L, self.Z[l][t], self._ctr[t], nh)
      else:
        off, n, k = self._w_off[(l, t)]
        pl.add_wgrad(self.dPre[l][t], self.A[l][t], self.g32[off:off + n * k].view(n, k), self._ctr[t], nh)
        if l > 1:
          pl.add_dgrad(self.dPre[l][t], self.W(l, t), self.dA[l][t], self._ctr[t], nh)
      self._plans[key] = pl
    return pl

  def _forward(self):
    nat, ar = self.nat, self.arena
    for l in range(1, self.L + 1):
      nh = self.L - l + 1
      for t in self.targets[l]:
        offs, self_col = self.col[l][t]
        A = self.A[l][t]
        for j, r in enumerate(self.in_rel[l][t]):
          s = self.nt[r]
          d = self._width_in(l, s)
          ell, ks = self._rel_ell(r, nh)
          if l == 1:
            nat.sage_aggregate_block(self.feat[self.ntypes[s]], ar.nodes_of(s), None, d, self._ctr[t], nh, ell, ks,
                                     ar.deg_of(r), A, offs[j])
          else:
            nat.sage_aggregate_block(None, None, self.Z[l - 1][s], d, self._ctr[t], nh, ell, ks, ar.deg_of(r), A,
                                     offs[j])
          self._k()
        d = self._width_in(l, t)
        blk = A[:, self_col:self_col + d]
        if l == 1:
          n = min(A.shape[0], self.cap_nodes[t])
          self.feat[self.ntypes[t]].gather_into(ar.nodes_of(t)[:n], None, self._ctr[t][nh:nh + 1], blk)
        else:
          blk.copy_(self.Z[l - 1][t][:A.shape[0]])
        self._k()
        self._plan('fwd', l, t).run()
        self._k()
    st = self.tid[self.seed_type]
    if self.peer_group is not None:
      self.peer_group.barrier(1)            # peers finished reading last step's gradients
      self._k()
    nat.zero_grads(self.g32, self.loss, self.correct)
    boff, n = self._b_off[(self.L, st)]
    nat.softmax_nll(self.Z[self.L][st], self.C, None, self.labels, ar.nodes_of(st), self._ctr[st], self.loss,
                    self.dPre[self.L][st], self.correct, self.g32[boff:boff + n], True)
    self._k(2)

  def _backward(self):
    nat, ar = self.nat, self.arena
    for l in range(self.L, 0, -1):
      nh = self.L - l + 1
      for t in self.targets[l]:
        self._plan('bwd', l, t).run()
        self._k()
      if l == 1:
        break
      H = self.hidden
      for s in self.targets[l - 1]:
        nat.zero_rows(self.dH[l - 1][s], self._ctr[s], nh + 1)
        self._k()
      for t in self.targets[l]:
        offs, self_col = self.col[l][t]
        for j, r in enumerate(self.in_rel[l][t]):
          ell, ks = self._rel_ell(r, nh)
          nat.sage_scatter_block(self.dA[l][t], H, offs[j], self._ctr[t], nh, ell, ks, ar.deg_of(r),
                                 self.dH[l - 1][self.nt[r]])
          self._k()
        nat.add_block_f32(self.dA[l][t], self_col, H, self._ctr[t], nh, self.dH[l - 1][t])
        self._k()
      for s in self.targets[l - 1]:
        boff, n = self._b_off[(l - 1, s)]
        nat.relu_bwd_cast(self.dH[l - 1][s], self.Z[l - 1][s], self._ctr[s], nh + 1, self.dPre[l - 1][s],
                          self.g32[boff:boff + n], True)
        self._k()

  def _optimizer(self):
    if self.peer_group is not None:
      self.peer_group.barrier(0)            # every rank finished writing its gradients
      self.peer_group.adam(self.p32, self.m, self.v, self.p16, self.lr, 0.9, 0.999, 1e-8, self.wd, self.step_dev,
                           1.0 / self.world)
      self._k(2)
      return
    self.nat.adam_step(self.p32, self.g32, self.m, self.v, self.p16, self.lr, 0.9, 0.999, 1e-8, self.wd,
                       self.step_dev, 1.0)
    self._k()

  def _step_eager(self):
    self._sample()
    self._forward()
    self._backward()
    self._optimizer()

  # ------------------------------------------------------------------ public API
  def warmup_and_capture(self, n_eager: int = 2):
    with torch.cuda.device(self.device):
      saved = self.state_dict()
      side = torch.cuda.Stream()
      side.wait_stream(torch.cuda.current_stream())
      with torch.cuda.stream(side):
        for _ in range(n_eager):
          self._tally = 0
          self._step_eager()
          self.kernels_per_step = self._tally
      torch.cuda.current_stream().wait_stream(side)
      torch.cuda.synchronize()
      self.load_state_dict(saved)
      if not self.use_cuda_graph:
        return
      g = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g):
        self._step_eager()
      torch.cuda.synchronize()
      self.load_state_dict(saved)
      self._graph = g

  def train_step(self, seeds: torch.Tensor) -> torch.Tensor:
    """One training step on a batch of seed ids of `seed_type` (host-pinned or device tensor); returns the device
    scalar holding the mean NLL loss."""
    n = seeds.numel()
    assert n <= self.bs
    if n < self.bs:
      self._seeds.fill_(-1)
    self._seeds[:n].copy_(seeds, non_blocking=True)
    if self._graph is not None:
      self._graph.replay()
    else:
      self._step_eager()
    self.step_idx += 1
    return self.loss

  @torch.no_grad()
  def evaluate_batch(self, seeds: torch.Tensor):
    n = seeds.numel()
    self._seeds.fill_(-1)
    self._seeds[:n].copy_(seeds)
    assert self.peer_group is None, 'evaluate_batch runs the gradient protocol of the multi-GPU step; use a single-GPU replica'
    self._sample()
    self._forward()
    st = self.tid[self.seed_type]
    return float(self.loss.item()), int(self.correct.item()), int(self._ctr[st][1].item())

  def overflow_count(self) -> int:
    return int(self.arena.counters[self._ovf_idx].item())

  def batch_sizes(self):
    """{type: nodes per hop} and {etype: edges per hop} of the last batch (host sync; diagnostics)."""
    c = self.arena.counters.cpu().tolist()
    NT = len(self.ntypes)
    nodes = {t: [c[i * 8 + h + 1] - c[i * 8 + h] for h in range(self.L + 1)] for i, t in enumerate(self.ntypes)}
    edges = {et: c[NT * 8 + r * 4:NT * 8 + r * 4 + self.L] for r, et in enumerate(self.edge_types)}
    return nodes, edges

  def close(self):
    """Drop the captured graph (call before destroying the process group)."""
    torch.cuda.synchronize(self.device)
    self._graph = None
