"""GraphSAGE: a generic PyG-style module and the fused, CUDA-graphed training engine.

The reference ships no model code; its examples use PyG's SAGEConv
(examples/train_sage_ogbn_products.py:30-59: 3 x SAGEConv, hidden 256, mean aggregation,
fanout [15,10,5], batch 1024).  `GraphSAGE` below is the drop-in eager module for the
`NeighborLoader` path.  `GraphSageEngine` is the B200-first path:

  sample (static-shape arena, no host sync)
    -> layer 1: ONE tcgen05 kernel = peer/local HBM row gather + neighbour mean +
       [mean | self] x W1 + bias + ReLU (csrc/cuda/sage_tc.cu)
    -> layers 2..L: ELL aggregation kernel + cuBLAS GEMM (small)
    -> fused log-softmax/NLL fwd+bwd -> hand-written backward -> NCCL all-reduce -> fused Adam

with every extent read from device counters, so the whole step is a fixed launch sequence
replayed from a CUDA graph.
"""
import math
from typing import List, Optional

import torch

from ..utils.tracing import nvtx_range
import torch.nn as nn
import torch.nn.functional as F

from ..ops import require_native


# ----------------------------------------------------------------------------- eager modules
class SAGEConv(nn.Module):
  """out_i = W_l . mean_{j in N(i)} x_j + W_r . x_i + b   (PyG SAGEConv semantics)."""

  def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
    super().__init__()
    self.lin_l = nn.Linear(in_channels, out_channels, bias=bias)
    self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

  def forward(self, x, edge_index, num_dst: Optional[int] = None):
    x_src, x_dst = (x, x) if not isinstance(x, tuple) else x
    n_dst = num_dst if num_dst is not None else x_dst.shape[0]
    x_dst = x_dst[:n_dst]
    src, dst = edge_index[0], edge_index[1]
    agg = torch.zeros(n_dst, x_src.shape[1], dtype=x_src.dtype, device=x_src.device)
    agg.index_add_(0, dst, x_src[src])
    deg = torch.zeros(n_dst, dtype=torch.float32, device=x_src.device)       # fp32: bf16 counts exactly only to 256
    deg.index_add_(0, dst, torch.ones_like(dst, dtype=torch.float32))
    agg = agg * (1.0 / deg.clamp(min=1)).to(agg.dtype).unsqueeze(1)
    return self.lin_l(agg) + self.lin_r(x_dst)


class GraphSAGE(nn.Module):
  """Eager GraphSAGE for `NeighborLoader` batches; trims targets layer by layer using
  `num_sampled_nodes/edges` when given (the reference relies on PyG trim_to_layer,
  examples/train_sage_prod_with_trim.py:60-65)."""

  def __init__(self, in_channels: int, hidden_channels: int, out_channels: int, num_layers: int = 3,
               dropout: float = 0.0):
    super().__init__()
    dims = [in_channels] + [hidden_channels] * (num_layers - 1) + [out_channels]
    self.convs = nn.ModuleList([SAGEConv(dims[i], dims[i + 1]) for i in range(num_layers)])
    self.dropout = dropout

  def forward(self, x, edge_index, num_sampled_nodes=None, num_sampled_edges=None):
    L = len(self.convs)
    for i, conv in enumerate(self.convs):
      n_dst = None
      ei = edge_index
      if num_sampled_nodes is not None and num_sampled_edges is not None:
        keep_hops = L - i                       # hops whose edges are still needed
        n_e = int(sum(num_sampled_edges[:keep_hops]))
        n_dst = int(sum(num_sampled_nodes[:keep_hops]))
        ei = edge_index[:, :n_e]
      x = conv(x, ei, n_dst)
      if i < L - 1:
        x = F.relu(x)
        if self.dropout > 0:
          x = F.dropout(x, self.dropout, self.training)
    return x


# ----------------------------------------------------------------------------- engine
def _round_up(x, m):
  return (x + m - 1) // m * m


class GraphSageEngine(object):
  """Static-shape, CUDA-graph-captured GraphSAGE training step.

  Args:
    graph: `data.Graph` in a device mode (single shard or multi-GPU shards).
    feature_table: native RowTableHandle of bf16 rows indexed by global node id
      (`UnifiedTensor._table()` / `PartitionedFeature.table`); width = in_dim.
    labels: int64 [num_nodes] on the device.
    in_dim: feature width (multiple of 64 for the fused kernel).
    num_nodes: number of graph nodes (arena sizing).
    fanouts / batch_size / hidden / num_classes: model + sampling config.
    group: torch.distributed group for the gradient all-reduce (None = single process).
    use_fused: True = use the tcgen05 gather+GEMM kernel where the shape allows; 'auto' = time the
      fused and the unfused (aggregate kernel + cuBLAS) layer on real batches during warm-up and
      keep the faster one (remote-heavy multi-GPU runs can favour the unfused pair).
    use_cuda_graph: capture sample+forward+backward(+adam) into CUDA graphs.
  """

  def __init__(self, graph, feature_table, labels: torch.Tensor, in_dim: int, num_nodes: int,
               fanouts: List[int] = (15, 10, 5), batch_size: int = 1024, hidden: int = 256,
               num_classes: int = 47, lr: float = 3e-3, weight_decay: float = 0.0, seed: int = 0,
               device: Optional[torch.device] = None, group=None, use_fused: bool = True,
               use_cuda_graph: bool = True, calibration_seeds: Optional[torch.Tensor] = None,
               calibration_margin: float = 1.3, calibration_batches: int = 16, pipeline: bool = False,
               use_peer_allreduce: bool = True, use_gather_bwd: Optional[bool] = None,
               check_every: int = 64, auto_regrow: bool = True, feature_format: str = 'bf16',
               deterministic_sampling: bool = False, stage_remote_rows: Optional[bool] = None,
               dropout: float = 0.0):
    self.nat = require_native()
    # dropout on the hidden activations (the reference example trains with p = 0.5,
    # examples/train_sage_ogbn_products.py:58): one in-place Philox kernel per hidden layer in training steps,
    # mask keyed by (seed, layer, optimizer step, element); backward rescales inside relu_bwd_cast.
    assert 0.0 <= float(dropout) < 1.0
    self.dropout = float(dropout)
    # 'mxfp8': `feature_table` holds data.quantize_mxfp8 rows (uint8, in_dim + 16 bytes); the fused layer-1 kernel
    # de-quantises in its loaders (half the gather bytes), so the fused path is mandatory for that format
    assert feature_format in ('bf16', 'mxfp8')
    self.feature_format = feature_format
    if feature_format == 'mxfp8':
      assert int(in_dim) == 128, 'MXFP8 features: in_dim must be 128'
      use_fused = True
    self.pipeline = bool(pipeline)
    # Multi-GPU: the rows of a batch's nodes that live in PEER HBM are copied once per batch into a local buffer
    # indexed by local node id -- by a bulk gather kernel on the SAMPLING stream, one batch ahead of the training
    # step -- and the fused layer-1 kernel reads them from there (each is needed ~2.5x, and 256-byte random reads
    # issued from inside the GEMM kernel are bound by NVLink latency, not bandwidth).  True = stage when the feature
    # table has non-local parts; None/False = in-kernel peer loads (the measured default, see _build_buffers).
    self._stage_remote_opt = stage_remote_rows
    import os as _os3
    # measured on B200 (profiles/): running the weight-gradient GEMMs and the zero fills on an auxiliary stream LOSES 3 %
    # (0.245 vs 0.237 ms/step): the extra CTAs compete with the critical dgrad -> scatter -> cast chain.  Kept as an option.
    self.overlap_wgrad = _os3.environ.get('GLT_B200_OVERLAP_WGRAD', '0') != '0'
    self.deterministic_sampling = bool(deterministic_sampling)
    self.use_peer_allreduce = bool(use_peer_allreduce)
    # use_gather_bwd (csrc/cuda/transpose.cu): the sampler also builds the transposed adjacency of the batch (on the
    # sampling stream) and the backward of the aggregation is an atomics-free gather (replaces zero_rows + fp32-atomic
    # scatter + relu_bwd_cast: 4 launches per hidden layer -> 2).  Measured on B200 with the v3 kernel: 0.2244-0.2264
    # vs 0.2363-0.2395 ms/step at products shape -> default on (GLT_B200_GATHER_BWD=0 restores the scatter path).
    if use_gather_bwd is None:
      import os as _os
      use_gather_bwd = _os.environ.get('GLT_B200_GATHER_BWD', '1') == '1'
    self.use_gather_bwd = bool(use_gather_bwd) and int(hidden) <= 1024
    # every dense contraction outside the fused layer-1 kernel runs on the TMA-fed tcgen05 GEMM kernel
    # (csrc/cuda/tc_gemm.cu); GLT_B200_TC_GEMM=0 falls back to cuBLAS for A/B measurements
    import os as _os2
    self.use_tc_gemm = _os2.environ.get('GLT_B200_TC_GEMM', '1') != '0'
    self._tc_plans = {}
    self.library_gemms_per_step = 0
    self.peer_group = None
    self.graph = graph
    graph.lazy_init()
    self.gh = graph.graph_handler
    self.feat = feature_table
    self.device = torch.device(device) if device is not None else torch.device('cuda', graph.device)
    self.labels = labels.to(self.device)
    self.fanouts = [int(k) for k in fanouts]
    self.L = len(self.fanouts)
    assert 1 <= self.L <= 4
    self.bs = int(batch_size)
    self.in_dim, self.hidden, self.C = int(in_dim), int(hidden), int(num_classes)
    assert self.in_dim % 8 == 0 and self.hidden % 8 == 0
    self.n_out_pad = _round_up(self.C, 64)
    self.lr, self.wd = lr, weight_decay
    self.seed = int(seed)
    self.group = group
    import torch.distributed as dist
    self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    self.use_cuda_graph = use_cuda_graph
    self.step_idx = 0
    self.kernels_per_step = 0

    dev = self.device
    self._num_nodes = int(num_nodes)
    # ---- health monitoring (peer-barrier timeouts, arena overflow) without stalling the step: every
    # `check_every` steps a tiny probe is copied to pinned memory and read one period later
    self.check_every = int(check_every)
    self.auto_regrow = bool(auto_regrow)
    self.regrow_count = 0
    self._overflow_seen = 0
    self._probe_event = None
    with torch.cuda.device(dev):
      cap_override = []
      if calibration_seeds is not None and calibration_seeds.numel() >= self.bs:
        cap_override = self._calibrate(calibration_seeds, int(num_nodes), calibration_margin, calibration_batches)
      self.calibrated = bool(cap_override)
      self.dims_in = [self.in_dim] + [self.hidden] * (self.L - 1)
      self.dims_out = [self.hidden] * (self.L - 1) + [self.n_out_pad]
      self._build_buffers(cap_override)
      self._init_params()
      n_arenas = len(self._arenas)
      self._seeds = [torch.zeros(self.bs, dtype=torch.int64, device=dev) for _ in range(n_arenas)]
      self._side = torch.cuda.Stream(device=dev) if self.pipeline else None
      self._aux = torch.cuda.Stream(device=dev)      # weight gradients + zero fills, off the critical chain
      self._primed = False
      self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
      self.correct = torch.zeros(1, dtype=torch.int32, device=dev)
      self.step_dev = torch.zeros(2, dtype=torch.int32, device=dev)   # {steps taken, block ticket}
      self._health_dev = torch.zeros(2, dtype=torch.int32, device=dev)
      self._health_host = torch.zeros(2, dtype=torch.int32).pin_memory()
      self._autotune_fused = (use_fused == 'auto')
      self.fused_ok = [False] + [bool(use_fused and self.nat.sage_fused_supported(self.dims_in[l - 1],
                                                                                    self.dims_out[l - 1]))
                                 for l in range(1, self.L + 1)]
      bf = torch.bfloat16
      self.w_packed = [None] + [torch.zeros(self.dims_out[l - 1] * 2 * self.dims_in[l - 1], dtype=bf, device=dev)
                                if self.fused_ok[l] else None for l in range(1, self.L + 1)]
      self._repack()
      if self.feature_format == 'mxfp8' and not self.fused_ok[1]:
        raise RuntimeError('MXFP8 features need the fused layer-1 kernel (in_dim 128, hidden multiple of 32, <= 256)')
    self._graph_fb = None
    self._graph_opt = None
    self._graph_full = None

  def _build_buffers(self, cap_override):
    """(Re)creates everything whose size depends on the arena capacities: sampler arenas, activation /
    gradient buffers, cached GEMM launches.  Parameters and optimizer state are untouched."""
    dev = self.device
    bf, f32 = torch.bfloat16, torch.float32
    # pipelined mode double-buffers the sampler state: batch i+1 is sampled on a side stream
    # while batch i trains (the reference's sampler<->trainer producer/consumer pipeline,
    # distributed/dist_sampling_producer.py:54-163, collapsed onto CUDA streams of one GPU)
    n_arenas = 2 if self.pipeline else 1
    self._arenas = [self.nat.SamplerArena(dev.index, self.bs, self.fanouts, False, self._num_nodes,
                                          list(cap_override) if cap_override else [])
                    for _ in range(n_arenas)]
    self._cur = 0
    for p_, ar_ in enumerate(self._arenas):
      ar_.step.fill_(p_ - n_arenas)          # disjoint Philox stream ids per arena
      ar_.deterministic = self.deterministic_sampling   # local ids of a hop in ascending global-id order
      if self.use_gather_bwd and self.L >= 2:
        ar_.enable_transpose(self.L - 1)     # layer 2 uses hops 0..L-2, deeper layers a prefix of them
    self.cap_rows = list(self.arena.cap_rows)           # frontier capacity per hop (+ last-hop additions)
    # layer l (1-based) targets = nodes of hops 0..L-l
    self.cap_T = [0] + [int(sum(self.cap_rows[:self.L - l + 1])) for l in range(1, self.L + 1)]
    self.cap_T = [min(c, self.arena.cap_nodes) for c in self.cap_T]
    self.A = [None] + [torch.zeros(self.cap_T[l], 2 * self.dims_in[l - 1], dtype=bf, device=dev)
                       for l in range(1, self.L + 1)]
    self.Z = [None] + [torch.zeros(self.cap_T[l], self.dims_out[l - 1], dtype=bf, device=dev)
                       for l in range(1, self.L + 1)]
    self.dPre = [None] + [torch.zeros(self.cap_T[l], self.dims_out[l - 1], dtype=bf, device=dev)
                          for l in range(1, self.L + 1)]
    self.dA = [None, None] + [torch.zeros(self.cap_T[l], 2 * self.dims_in[l - 1], dtype=bf, device=dev)
                              for l in range(2, self.L + 1)]
    self.dH = [None] + [torch.zeros(self.cap_T[l], self.dims_out[l - 1], dtype=f32, device=dev)
                        for l in range(1, self.L)]
    want = self._stage_remote_opt
    if want is None:
      import os as _os4
      # measured (profiles/r2_gpu_call7_8gpu_session.log, call11): staging is neutral at 2 GPUs (0.275 vs 0.270 ms)
      # and LOSES 4 % at 8 GPUs (0.308 vs 0.296 ms/step) -- the staging kernel competes with the sampler for the
      # side stream and the unfused layer 1 already hides the NVLink latency behind its occupancy -> opt-in
      want = _os4.environ.get('GLT_B200_STAGE_REMOTE', '0') != '0'
    has_remote = hasattr(self.feat, 'all_local') and not self.feat.all_local()
    self.stage_remote = bool(want and has_remote)
    self._xcache = None
    if self.stage_remote:
      row_bytes = self.in_dim * 2 if self.feature_format == 'bf16' else self.in_dim + 16
      self._xcache = [torch.zeros(int(a_.cap_nodes), row_bytes, dtype=torch.uint8, device=dev) for a_ in self._arenas]
    self._tc_plans = {}
    self._graphs = []
    self._graph_fb = self._graph_opt = self._graph_full = None

  def regrow(self, factor: float = 1.5):
    """Enlarge the (calibrated) arena after neighbours were dropped by the capacity guard: new per-hop
    capacities = old x factor (bounded by the worst case), buffers and CUDA graphs are rebuilt, parameters,
    optimizer state and the sampling position are kept.  Collective when world > 1 (every rank calls it at the
    same step -- `_health_tick` guarantees that)."""
    with torch.cuda.device(self.device):
      torch.cuda.synchronize()
      steps = [int(a.step.item()) for a in self._arenas]
      tmp = self.nat.SamplerArena(self.device.index, self.bs, self.fanouts, False, self._num_nodes)
      worst = list(tmp.cap_rows)
      del tmp
      caps = [self.bs] + [min(worst[h], (int(self.cap_rows[h] * factor) + 255) // 128 * 128)
                          for h in range(1, self.L + 1)]
      had_graphs = bool(getattr(self, '_graphs', None))
      self._build_buffers(caps)
      for a, v in zip(self._arenas, steps):
        a.step.fill_(v)
      self._primed = False
      self.regrow_count += 1
      if had_graphs:
        self.warmup_and_capture(n_eager=1)

  def _health_tick(self):
    """Called once per train_step.  Every `check_every` steps: (1) read the probe enqueued one period ago --
    raise if a peer barrier timed out (the Adam kernels already skipped that update), regrow the arena if
    neighbours were dropped; (2) enqueue a new probe (max-reduced over ranks so that every rank takes the same
    decision at the same step)."""
    if self.check_every <= 0 or self.step_idx % self.check_every != 0:
      return
    if self._probe_event is not None:
      self._probe_event.synchronize()
      err, ovf = int(self._health_host[0]), int(self._health_host[1])
      self._probe_event = None
      if err != 0:
        raise RuntimeError(f'peer barrier timed out waiting for rank {err - 1}: a peer process died or stalled; '
                           'the optimizer updates since then were skipped')
      if ovf > self._overflow_seen:
        self._overflow_seen = ovf
        if self.auto_regrow:
          self.regrow()
    h = self._health_dev
    h.zero_()
    for pg in self._peer_groups:
      torch.maximum(h[0:1], pg.err, out=h[0:1])
    for a in self._arenas:
      h[1:2].add_(a.counters[12:13])
    if self.regrow_count:
      h[1:2].add_(self._overflow_seen)      # counters restart from zero in a rebuilt arena
    if self.world > 1:
      import torch.distributed as dist
      dist.all_reduce(h, op=dist.ReduceOp.MAX, group=self.group)
    self._health_host.copy_(h, non_blocking=True)
    self._probe_event = torch.cuda.Event()
    self._probe_event.record()

  @property
  def arena(self):
    return self._arenas[self._cur]

  @property
  def seeds_dev(self):
    return self._seeds[self._cur]

  @property
  def g32(self):
    """Flat fp32 gradient buffer of the batch being trained (per parity when double-buffered)."""
    return self._g[self._cur % len(self._g)]

  # ------------------------------------------------------------------ capacity calibration
  def _calibrate(self, pool: torch.Tensor, num_nodes: int, margin: float, n_batches: int):
    """Size the arena from observed frontier sizes instead of the worst case
    bs*prod(fanouts): sample a few batches of the seed pool with a worst-case arena, take the
    per-hop maxima (max over ranks), add a margin.  All buffers and every cap-proportional
    launch (cuBLAS GEMMs, fills) shrink accordingly.  The kernels stay safe past the caps
    (excess nodes are dropped and counted in `overflow_count()`)."""
    dev = self.device
    tmp = self.nat.SamplerArena(dev.index, self.bs, self.fanouts, False, num_nodes)
    worst = list(tmp.cap_rows)
    mx = torch.zeros(self.L + 1, dtype=torch.int64, device=dev)
    g = torch.Generator(device='cpu')
    g.manual_seed(self.seed + 991)
    for i in range(n_batches):
      idx = torch.randint(0, pool.numel(), (self.bs,), generator=g)
      seeds = pool[idx].to(dev)
      tmp.sample(self.gh, seeds.contiguous(), None, self.seed + 17, i * 8, False, False, False)
      c = tmp.counters[:self.L + 2].to(torch.int64)
      mx = torch.maximum(mx, c[1:] - c[:-1])
    if self.world > 1:
      import torch.distributed as dist
      dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
    mx = mx.cpu().tolist()
    caps = []
    for h in range(self.L + 1):
      c = int(mx[h] * margin) + 128
      c = (c + 127) // 128 * 128
      caps.append(min(c, worst[h]) if h > 0 else self.bs)
    del tmp
    torch.cuda.empty_cache()
    return caps

  def overflow_count(self) -> int:
    """Neighbours dropped by the capacity guard since construction (0 in normal operation)."""
    return int(sum(int(a.counters[12].item()) for a in self._arenas))

  # ------------------------------------------------------------------ parameters
  def _init_params(self):
    dev = self.device
    sizes = []
    for l in range(self.L):
      sizes.append(self.dims_out[l] * 2 * self.dims_in[l])   # W [N, 2d] = [W_neigh | W_self]
      sizes.append(self.dims_out[l])                          # bias
    total = sum(sizes)
    self.p32 = torch.zeros(total, dtype=torch.float32, device=dev)
    g = torch.Generator(device='cpu')
    g.manual_seed(self.seed)
    off = 0
    self.W32, self.b32, self._w_off, self._b_off = [], [], [], []
    for l in range(self.L):
      n, k = self.dims_out[l], 2 * self.dims_in[l]
      bound = 1.0 / math.sqrt(self.dims_in[l])
      w = (torch.rand(n, k, generator=g) * 2 - 1) * bound
      if l == self.L - 1 and self.n_out_pad > self.C:
        w[self.C:] = 0
      self.p32[off:off + n * k].copy_(w.flatten())
      self._w_off.append((off, n, k)); off += n * k
      self._b_off.append((off, n)); off += n
    self.p16 = self.p32.to(torch.bfloat16)
    self._g = [torch.zeros_like(self.p32)]
    self._peer_groups = []
    if self.world > 1 and self.use_peer_allreduce:
      # gradients live in a symmetric-heap segment: the optimizer kernel of every rank reads
      # all peers' gradients over NVLink (all-reduce fused into Adam, no NCCL in the step)
      import torch.distributed as dist
      from ..parallel.peer import exchange_peer_tensors
      rank = dist.get_rank(self.group)
      # pipelined engines run parity graphs, so each parity gets its own gradient segment and
      # flag set: one cross-GPU barrier per step is then enough (the buffer written at step k is
      # not touched again before step k+2, after everybody passed the barrier of step k+1)
      n_par = 2 if self.pipeline else 1
      self._g = []
      for _ in range(n_par):
        peers_g = exchange_peer_tensors(torch.zeros_like(self.p32), self.group)
        flags = exchange_peer_tensors(torch.zeros(2 * self.world, dtype=torch.int32, device=dev), self.group)
        self._g.append(peers_g[rank])
        self._peer_groups.append(self.nat.PeerGroup(dev.index, rank, peers_g, flags))
      self.peer_group = self._peer_groups[0]
    self.m = torch.zeros_like(self.p32)
    self.v = torch.zeros_like(self.p32)

  def W(self, l, buf=None):      # l is 1-based
    off, n, k = self._w_off[l - 1]
    return (buf if buf is not None else self.p16)[off:off + n * k].view(n, k)

  def b(self, l, buf=None):
    off, n = self._b_off[l - 1]
    return (buf if buf is not None else self.p16)[off:off + n]

  def _repack(self):
    for l in range(1, self.L + 1):
      if self.fused_ok[l]:
        self.nat.pack_weight_into(self.W(l), self.w_packed[l])
        self._k(1)

  def state_dict(self):
    return {'p32': self.p32.clone(), 'm': self.m.clone(), 'v': self.v.clone(),
            'step': int(self.step_dev[0].item()), 'sample_step': [int(a.step.item()) for a in self._arenas],
            'step_idx': self.step_idx, 'seed': self.seed}

  def load_state_dict(self, s):
    self.p32.copy_(s['p32']); self.m.copy_(s['m']); self.v.copy_(s['v'])
    self.p16.copy_(self.p32)
    self.step_dev.zero_()
    self.step_dev[0] = int(s['step'])
    ss = s.get('sample_step', None)
    if ss is not None:
      for a, v in zip(self._arenas, ss if isinstance(ss, (list, tuple)) else [ss]):
        a.step.fill_(v)
    self.step_idx, self.seed = s['step_idx'], s['seed']
    self._repack()

  # ------------------------------------------------------------------ step pieces
  def _k(self, n):
    self._tally += n

  def _lib(self, n):
    """Counts launches that go to a vendor library (cuBLAS/cuBLASLt) instead of this repo's kernels."""
    self._lib_tally += n

  _tally = 0
  _lib_tally = 0

  def _ell(self, l):
    """ELL blocks + strides used by layer l: hops 0..L-l."""
    nh = self.L - l + 1
    return list(self.arena.ell[:nh]), self.fanouts[:nh], nh

  def _sample(self, which: Optional[int] = None):
    # the Philox stream advances from a device-side step counter (arena.step) so that a
    # CUDA-graph replay draws fresh samples every step
    which = self._cur if which is None else which
    ar = self._arenas[which]
    # (the increment itself is folded into the first sampling kernel: step_inc)
    ar.sample(self.gh, self._seeds[which], None, self.seed, 0, False, False, True, len(self._arenas))
    self._k(2 + 2 * self.L)  # table clear + init_seeds + (sample, relabel) per hop
    if self.stage_remote:
      self.feat.stage_remote_rows(ar.nodes, ar.counters, self.L + 1, self._xcache[which])
      self._k(1)

  def _forward(self, train: bool = True):
    nat, ar = self.nat, self.arena
    for l in range(1, self.L + 1):
      self._forward_layer(l, train)
    self._forward_loss(train)

  def _forward_layer(self, l: int, train: bool = True):
    self._forward_layer_nodrop(l)
    if train and self.dropout > 0.0 and l < self.L:
      self.nat.dropout_bf16(self.Z[l], self.arena.counters, self.L - l + 1, self.dropout, self.seed, l,
                            self.step_dev)
      self._k(1)

  def _forward_layer_nodrop(self, l: int):
    nat, ar = self.nat, self.arena
    ell, ks, nh = self._ell(l)
    d = self.dims_in[l - 1]
    relu = l < self.L
    feat = self.feat if l == 1 else None
    nodes = ar.nodes if l == 1 else None
    src_local = None if l == 1 else self.Z[l - 1]
    if self.fused_ok[l]:
      xc = self._xcache[self._cur] if (l == 1 and self.stage_remote) else None
      nat.sage_fused(feat, nodes, src_local, d, ar.counters, nh, ell, ks, ar.deg, self.w_packed[l],
                     self.b(l), relu, self.Z[l], self.A[l], xc)
      self._k(1)
    else:
      nat.sage_aggregate(feat, nodes, src_local, d, ar.counters, nh, ell, ks, ar.deg, self.A[l])
      self._k(1)
      if self.use_tc_gemm:
        # Z = act(A . W^T + b) on tcgen05 (TMA-fed, bias/ReLU in the TMEM epilogue, rows from the device counter)
        self._plan('fwd', l).run()
        self._k(1)
      else:
        # library GEMM with the bias (+ ReLU) applied in the cuBLASLt epilogue
        if relu:
          torch._addmm_activation(self.b(l), self.A[l], self.W(l).t(), out=self.Z[l])
        else:
          torch.addmm(self.b(l), self.A[l], self.W(l).t(), out=self.Z[l])
        self._lib(1)

  def _forward_loss(self, train: bool = True):
    # labels[nodes[r]] is looked up inside the loss kernel (no gather launch)
    nat, ar = self.nat, self.arena
    boff, n = self._b_off[self.L - 1]
    if train:                      # evaluation must not take part in the cross-rank gradient protocol
      self._begin_grads()
    # the bias gradient of the last layer (column sums of dlogits) is produced by the loss kernel
    nat.softmax_nll(self.Z[self.L], self.C, None, self.labels, ar.nodes, ar.counters, self.loss,
                    self.dPre[self.L], self.correct, self.g32[boff:boff + n], train)
    self._k(1)

  def _begin_grads(self):
    """Runs before the first kernel that writes into the flat gradient buffer (the loss kernel produces the
    last layer's bias gradient): peers must be done reading last step's gradients, and the split-K weight
    gradient kernels accumulate with red.add, so the buffer starts from zero."""
    if self.peer_group is not None and len(self._peer_groups) == 1:
      self.peer_group.barrier(1)          # peers finished reading last step's gradients
      self._k(1)
    # one launch zeroes the flat gradient buffer, the loss and the #correct counter; the kernels that accumulate
    # into them then skip their own memsets (memset nodes would cut the programmatic-launch chain of the step)
    self.nat.zero_grads(self.g32, self.loss, self.correct)
    self._k(1)

  def _plan(self, kind: str, l: int):
    """Cached TcGemm launch for layer l of the current arena / gradient parity (tensor maps are encoded once)."""
    key = (kind, l, self._cur)
    pl = self._tc_plans.get(key)
    if pl is not None:
      return pl
    nat, ar = self.nat, self.arena
    nh = self.L - l + 1                     # counters[nh] = number of target rows of layer l
    pl = nat.TcGemm(self.device.index)
    if kind == 'fwd':
      pl.add_forward(self.A[l], self.W(l), self.b(l), l < self.L, self.Z[l], ar.counters, nh)
    else:
      off, n, k = self._w_off[l - 1]
      gW = self.g32[off:off + n * k].view(n, k)
      # 'bwd': dW (split-K over the batch rows, fp32 red-add) and dA of the same layer share one launch;
      # 'wgrad' / 'dgrad': the two halves as separate launches (overlap_wgrad: dW leaves the critical chain)
      if kind in ('bwd', 'wgrad'):
        pl.add_wgrad(self.dPre[l], self.A[l], gW, ar.counters, nh)
      if l > 1 and kind in ('bwd', 'dgrad'):
        pl.add_dgrad(self.dPre[l], self.W(l), self.dA[l], ar.counters, nh)
    self._tc_plans[key] = pl
    return pl

  def _backward(self):
    nat, ar = self.nat, self.arena
    overlap = self.overlap_wgrad and self.use_tc_gemm and not self.use_gather_bwd and self.L > 1
    main = torch.cuda.current_stream()
    if overlap:
      # Only dA -> scatter -> ReLU-mask feeds the next layer; the weight gradients are consumed by Adam alone.  They
      # run on an auxiliary stream (forked after dPre[l] exists, joined before the optimizer) together with the
      # zero-fill of the fp32 scatter targets, so the critical chain per layer is dgrad -> scatter -> cast.
      aux = self._aux
      aux.wait_stream(main)
      with torch.cuda.stream(aux):
        for l in range(self.L, 1, -1):
          nat.zero_rows(self.dH[l - 1], ar.counters, self.L - l + 2)
        self._k(self.L - 1)
      ev_zero = torch.cuda.Event()
      ev_zero.record(aux)
    for l in range(self.L, 0, -1):
      ell, ks, nh = self._ell(l)
      off, n, k = self._w_off[l - 1]
      boff, _ = self._b_off[l - 1]
      if overlap and l > 1:
        ev = torch.cuda.Event()
        ev.record(main)                     # dPre[l] (and everything before it) is complete
        with torch.cuda.stream(aux):
          aux.wait_event(ev)
          self._plan('wgrad', l).run()
        self._plan('dgrad', l).run()
        self._k(2)
        pboff, pn = self._b_off[l - 2]
        if l == self.L:
          main.wait_event(ev_zero)
        nat.sage_scatter_bwd(self.dA[l], self.dims_in[l - 1], ar.counters, nh, ell, ks, ar.deg, self.dH[l - 1])
        nat.relu_bwd_cast(self.dH[l - 1], self.Z[l - 1], ar.counters, nh + 1, self.dPre[l - 1],
                          self.g32[pboff:pboff + pn], True, 1.0 / (1.0 - self.dropout))
        self._k(2)
        continue
      if self.use_tc_gemm:
        self._plan('wgrad' if overlap else 'bwd', l).run()
        self._k(1)
        if overlap:
          main.wait_stream(aux)             # join: every weight gradient is in the flat buffer
      else:
        gW = self.g32[off:off + n * k].view(n, k)
        # dW = dPre^T A accumulated in fp32 straight into the flat gradient buffer
        self._mm_f32(self.dPre[l].t(), self.A[l], gW)
        self._lib(1)
        if l > 1:
          torch.mm(self.dPre[l], self.W(l), out=self.dA[l])
          self._lib(1)
      # bias gradients are fused into the kernels that produce dPre (loss / relu_bwd_cast)
      if l > 1:
        pboff, pn = self._b_off[l - 2]
        if self.use_gather_bwd:
          nat.sage_gather_bwd(self.dA[l], self.dims_in[l - 1], ar, nh, self.Z[l - 1], self.dPre[l - 1],
                              self.g32[pboff:pboff + pn], True, 1.0 / (1.0 - self.dropout))
          self._k(1)
          continue
        nat.zero_rows(self.dH[l - 1], ar.counters, nh + 1)
        self._k(1)
        nat.sage_scatter_bwd(self.dA[l], self.dims_in[l - 1], ar.counters, nh, ell, ks, ar.deg, self.dH[l - 1])
        nat.relu_bwd_cast(self.dH[l - 1], self.Z[l - 1], ar.counters, nh + 1, self.dPre[l - 1],
                          self.g32[pboff:pboff + pn], True, 1.0 / (1.0 - self.dropout))
        self._k(2)

  _f32_mode = None

  def _mm_f32(self, a, b, out):
    """out(fp32) = a(bf16) @ b(bf16) with fp32 accumulation/output when cuBLASLt offers it."""
    if GraphSageEngine._f32_mode is None:
      try:
        torch.mm(a, b, out_dtype=torch.float32, out=out)
        GraphSageEngine._f32_mode = 'out'
        return
      except Exception:
        try:
          out.copy_(torch.mm(a, b, out_dtype=torch.float32))
          GraphSageEngine._f32_mode = 'ret'
          return
        except Exception:
          GraphSageEngine._f32_mode = 'bf16'
    if GraphSageEngine._f32_mode == 'out':
      torch.mm(a, b, out_dtype=torch.float32, out=out)
    elif GraphSageEngine._f32_mode == 'ret':
      out.copy_(torch.mm(a, b, out_dtype=torch.float32))
    else:
      out.copy_(torch.mm(a, b))

  def _optimizer(self):
    # Adam's step counter is advanced by the kernel itself (last block to finish)
    if self.peer_group is not None:
      pg = self._peer_groups[self._cur % len(self._peer_groups)]
      pg.barrier(0)                       # every rank finished writing its gradients
      pg.adam(self.p32, self.m, self.v, self.p16, self.lr, 0.9, 0.999, 1e-8, self.wd,
                           self.step_dev, 1.0 / self.world)
      self._k(2)
    else:
      self.nat.adam_step(self.p32, self.g32, self.m, self.v, self.p16, self.lr, 0.9, 0.999, 1e-8, self.wd,
                         self.step_dev, 1.0 / self.world)
      self._k(1)
    self._repack()

  def _allreduce(self):
    if self.world > 1 and self.peer_group is None:
      import torch.distributed as dist
      dist.all_reduce(self.g32, group=self.group)

  def _pipelined_body(self, cur: int):
    """train(batch in arena[cur])  ||  sample(next batch into arena[1-cur])."""
    self._cur = cur
    main = torch.cuda.current_stream()
    self._side.wait_stream(main)                       # fork
    with torch.cuda.stream(self._side):
      self._sample(1 - cur)
    self._forward()
    self._backward()
    main.wait_stream(self._side)                        # join

  def _step_eager(self):
    if self.pipeline:
      self._pipelined_body(self._cur)
    else:
      self._sample()
      self._forward()
      self._backward()
    self._allreduce()
    self._optimizer()

  def _autotune(self, iters: int = 5):
    """Pick fused vs unfused per layer from device timings on the current batch."""
    self.autotune_ms = {}
    ar = self.arena
    for l in range(1, self.L + 1):
      if not self.fused_ok[l]:
        continue
      res = {}
      for mode in (True, False):
        self.fused_ok[l] = mode
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self._forward_layer_nodrop(l)
        e0.record()
        for _ in range(iters):
          self._forward_layer_nodrop(l)
        e1.record()
        e1.synchronize()
        res[mode] = e0.elapsed_time(e1) / iters
      t = torch.tensor([res[True], res[False]], device=self.device)
      if self.world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)   # same choice on every rank
      self.fused_ok[l] = bool(t[0] <= t[1])
      self.autotune_ms[l] = {'fused': float(t[0]), 'unfused': float(t[1])}

  # ------------------------------------------------------------------ public API
  def warmup_and_capture(self, n_eager: int = 2):
    """Run a few eager steps (lazy inits, cuBLAS workspaces), then capture the CUDA graphs.
    Pipelined engines capture one graph per buffer parity."""
    import os
    with torch.cuda.device(self.device):
      saved = self.state_dict()
      side = torch.cuda.Stream()
      side.wait_stream(torch.cuda.current_stream())
      with torch.cuda.stream(side):
        if self.pipeline:
          self._sample(0)
        for _ in range(n_eager):
          self._tally = self._lib_tally = 0
          self._step_eager()
          self.kernels_per_step = self._tally
          self.library_gemms_per_step = self._lib_tally
          if self.pipeline:
            self._cur ^= 1
      torch.cuda.current_stream().wait_stream(side)
      torch.cuda.synchronize()
      if self._autotune_fused:
        self._cur = 0
        self._autotune()
        self._autotune_fused = False
        if self.pipeline:
          self._sample(0)
        self._tally = self._lib_tally = 0
        self._step_eager()                      # re-count kernels with the chosen variants
        self.kernels_per_step = self._tally
        self.library_gemms_per_step = self._lib_tally
        torch.cuda.synchronize()
        self._cur = 0
      self.load_state_dict(saved)
      self._cur, self._primed = 0, False
      self._graphs = []
      if not self.use_cuda_graph:
        return
      # programmatic dependent launch helps single-stream chains and costs ~2 % when the sampling and training
      # streams interleave (measured, csrc/cuda/launch_utils.h): captured without it in pipelined mode
      prev_pdl = self.nat.set_pdl(not self.pipeline) if hasattr(self.nat, 'set_pdl') else None
      # capturing NCCL collectives works but makes process-group teardown hang on this stack
      # (measured: bench exit blocked until the timeout), so it is opt-in for world > 1
      single = self.world == 1 or self.peer_group is not None or \
          os.environ.get('GLT_B200_CAPTURE_NCCL', '0') == '1'
      g_opt = None
      for parity in range(2 if self.pipeline else 1):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
          if self.pipeline:
            self._pipelined_body(parity)
          else:
            self._sample()
            self._forward()
            self._backward()
          if single:
            self._allreduce()
            self._optimizer()
        if not single and g_opt is None:
          g_opt = torch.cuda.CUDAGraph()
          with torch.cuda.graph(g_opt):
            self._optimizer()
        self._graphs.append((g, None if single else g_opt))
      torch.cuda.synchronize()
      if prev_pdl is not None:
        self.nat.set_pdl(prev_pdl)
      self.load_state_dict(saved)
      self._cur, self._primed = 0, False
      self._graph_fb = self._graphs[0][0]

  def close(self):
    """Drop the captured graphs (call before destroying the process group)."""
    torch.cuda.synchronize(self.device)
    self._graphs = []
    self._graph_full = self._graph_fb = self._graph_opt = None

  def _run_captured_or_eager(self, parity: int):
    if getattr(self, '_graphs', None):
      g, g_opt = self._graphs[parity]
      g.replay()
      if g_opt is not None:
        self._allreduce()
        g_opt.replay()
    else:
      self._cur = parity
      self._step_eager()

  def _stage_seeds(self, buf: torch.Tensor, seeds: torch.Tensor):
    n = seeds.numel()
    assert n <= self.bs
    if n < self.bs:
      buf.fill_(-1)
    buf[:n].copy_(seeds, non_blocking=True)

  def train_step(self, seeds: torch.Tensor) -> Optional[torch.Tensor]:
    """One training step on a batch of seed node ids.  `seeds` may be a (pinned) host
    tensor -- it is copied H2D asynchronously -- or a device tensor.  Returns the device
    scalar holding the mean NLL loss (read it with .item() / a D2H copy).

    Pipelined engines run *sample(this batch) || train(previous batch)*: the first call only
    samples (returns None), later calls return the loss of the batch passed one call earlier;
    `flush()` trains the last pending batch."""
    self._health_tick()
    if not self.pipeline:
      with nvtx_range('glt.engine.step'):
        self._stage_seeds(self._seeds[0], seeds)
        self._run_captured_or_eager(0)
      self.step_idx += 1
      return self.loss
    if not self._primed:
      self._cur = 0
      self._stage_seeds(self._seeds[0], seeds)
      self._sample(0)
      self._primed = True
      return None if self.step_idx == 0 and self.regrow_count == 0 else self.loss
    cur = self._cur
    with nvtx_range('glt.engine.step(sample b+1 || train b)'):
      self._stage_seeds(self._seeds[1 - cur], seeds)
      self._run_captured_or_eager(cur)
    self._cur = 1 - cur
    self.step_idx += 1
    return self.loss

  def flush(self) -> Optional[torch.Tensor]:
    """Pipelined mode: train the batch that was sampled by the last train_step() call."""
    if not self.pipeline or not self._primed:
      return None
    self._forward()
    self._backward()
    self._allreduce()
    self._optimizer()
    self._primed = False
    self.step_idx += 1
    return self.loss

  def profile_sections(self, seeds: torch.Tensor, iters: int = 10):
    """Device time (ms) of each stage of an *eager, unpipelined* step (CUDA events; the other
    ranks run the same sequence so the peer barriers inside are matched).  Diagnostic only."""
    names = ['sample', 'forward_l1', 'forward_rest', 'backward', 'optimizer']
    acc = {n: 0.0 for n in names}
    self._cur = 0
    # `seeds` may hold several batches ([n_batches, batch]): rotating them keeps the feature rows
    # of every iteration cold in L2, like a real epoch
    batches = seeds if seeds.dim() == 2 else seeds.unsqueeze(0)
    for it in range(iters + 2):
      self._stage_seeds(self._seeds[0], batches[it % batches.shape[0]])
      ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
      ev[0].record()
      self._sample(0); ev[1].record()
      self._forward_layer(1); ev[2].record()
      for l in range(2, self.L + 1):
        self._forward_layer(l)
      self._forward_loss(); ev[3].record()
      self._backward(); self._allreduce(); ev[4].record()
      self._optimizer(); ev[5].record()
      torch.cuda.synchronize()
      if it >= 2:
        for i, n in enumerate(names):
          acc[n] += ev[i].elapsed_time(ev[i + 1]) / iters
    return acc

  def profile_kernels(self, seeds: torch.Tensor, iters: int = 10, queue_ahead_cycles: int = 6_000_000):
    """Device time (us) of EVERY launch of an eager, unpipelined step, in launch order: -> [(name, us)].
    Each native call is bracketed by CUDA events; a spin kernel at the head of the step lets the host queue the
    whole step ahead of the GPU, so an interval is the kernel's own run time with the previous kernel's data still
    in L2 (what `ncu`, which replays each kernel with cold caches, cannot show).  Diagnostic only."""
    eng = self
    log = []

    class _Timed(object):
      def __init__(self, obj, prefix=''):
        object.__setattr__(self, '_o', obj)
        object.__setattr__(self, '_p', prefix)

      def __getattr__(self, name):
        f = getattr(self._o, name)
        if name in ('TcGemm', 'TcGemmMx'):
          return lambda *a, **k: _Timed(f(*a, **k), name + '.')
        if not callable(f) or name.startswith('_') or isinstance(f, type) or name in ('set_pdl',):
          return f

        def call(*a, **k):
          e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          e0.record()
          r = f(*a, **k)
          e1.record()
          log.append((self._p + name, e0, e1))
          return r
        return call

    real_nat, real_plans, real_pgs = self.nat, self._tc_plans, self._peer_groups
    batches = seeds if seeds.dim() == 2 else seeds.unsqueeze(0)
    acc, names = None, None
    try:
      self.nat, self._tc_plans = _Timed(real_nat), {}
      self._peer_groups = [_Timed(pg, 'PeerGroup.') for pg in real_pgs]
      if real_pgs:
        self.peer_group = self._peer_groups[0]
      self._cur = 0
      for it in range(iters + 2):
        del log[:]
        self._stage_seeds(self._seeds[0], batches[it % batches.shape[0]])
        torch.cuda._sleep(int(queue_ahead_cycles))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ar = self._arenas[0]
        ar.sample(self.gh, self._seeds[0], None, self.seed, 0, False, False, True, len(self._arenas))
        e1.record()
        log.append(('arena.sample (all sampling kernels)', e0, e1))
        if self.stage_remote:
          self.feat.stage_remote_rows(ar.nodes, ar.counters, self.L + 1, self._xcache[0])
        self._forward()
        self._backward()
        self._allreduce()
        self._optimizer()
        torch.cuda.synchronize()
        if it >= 2:
          if acc is None:
            names, acc = [n for n, _, _ in log], [0.0] * len(log)
          if len(log) == len(acc):
            for i, (_, a0, a1) in enumerate(log):
              acc[i] += a0.elapsed_time(a1) * 1e3 / iters
    finally:
      self.nat, self._tc_plans, self._peer_groups = real_nat, real_plans, real_pgs
      if real_pgs:
        self.peer_group = real_pgs[0]
    return list(zip(names, acc))

  @torch.no_grad()
  def evaluate_batch(self, seeds: torch.Tensor):
    """(loss, #correct, #seeds) for a batch without updating parameters."""
    assert not (self.pipeline and self._primed), 'call flush() before evaluate_batch()'
    n = seeds.numel()
    self.seeds_dev.fill_(-1)
    self.seeds_dev[:n].copy_(seeds)
    self._sample()
    self._forward(train=False)
    return float(self.loss.item()), int(self.correct.item()), int(self.arena.counters[1].item())
