#!/usr/bin/env python
"""Headline benchmark: GraphSAGE training throughput (seed nodes / s) on an
ogbn-products-shape synthetic graph (BASELINE.json metric/config).

  python bench.py --gpus N --steps K --warmup W              # this framework (engine path, bf16, HBM-resident)
  python bench.py --impl reference --gpus N ...               # unmodified reference (baseline/_ref), same pairing
  python bench.py --path loader --dtype fp32                  # this framework through NeighborLoader + eager model
  python bench.py --shape papers100m --gpus 8                 # BASELINE config 3 (111 M nodes / 1.6 B edges / F=128)

Model/config (reference examples/train_sage_ogbn_products.py:30-59,112-114): 3-layer GraphSAGE,
hidden 256, fanout [15,10,5], batch 1024 seeds per GPU, Adam, NLL loss; graph = RMAT with the
ogbn-products shape (2,449,029 nodes, 123.7 M directed edges, 100-dim features, 47 classes),
random-init weights, synthetic features/labels (no network access for the real dataset).

PAIRING (what is compared with what).  The default arms are like-for-like:

  ours      : graph + features fully in HBM, bf16 feature storage, bf16 tensor-core GEMMs, fp32 master weights
  reference : graph_mode='CUDA', split_ratio=1.0 (its own HBM-resident recipe,
              examples/train_sage_prod_with_trim.py:97,103), bf16 feature storage, bf16 autocast GEMMs,
              fp32 master weights

`--ref-config stock --ref-dtype fp32` reproduces the reference example's stock setting (ZERO_COPY topology,
20 % of the features in HBM, fp32).  Both arms also report secondary pairings under "arms": ours adds the
loader path (glt.loader.NeighborLoader + the SAME plain-PyTorch model code the reference arm uses) in fp32
and bf16; the reference adds its fp32/HBM number.

TIMING.  After W warm-up steps the K-step region is timed with CUDA events (barrier + synchronize on both
sides, max over ranks).  A single K-step region lasts only milliseconds here, so the region is repeated
back to back on FRESH seed batches until at least --min-time seconds of device time have been measured;
`ms_per_step` / `value` are the mean over all timed steps (`timed.steps_total`), `timed.first_block_ms_per_step`
is the first K-step block alone.  `e2e` re-measures the same metric through the public API call
(`GraphSageEngine.train_step(pinned_host_seeds)` / the loader iterator) including the per-step H2D copy of the
step's seed ids from pinned host memory and a D2H read of the step's loss.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# derived from BASELINE.md: GLT GraphSAGE papers100M epoch = 8.56 s on 2 nodes x 2 A100 with
# 1,207,179 training seeds  ->  141,025 seed nodes / s (the only published end-to-end training
# throughput of the reference; other hardware, other graph -- context only).
BASELINE_SAMPLES_PER_S = 1_207_179 / 8.56

SHAPES = {
  'products': dict(nodes=2_449_029, edges=123_718_280, feat_dim=100, classes=47),
  'papers100m': dict(nodes=111_059_956, edges=1_615_685_872, feat_dim=128, classes=172),
}
METRIC = 'GraphSAGE ogbn-products-shape training throughput (seed nodes/s, device-timed, max over ranks)'


def metric_name(args):
  return METRIC if args.shape == 'products' else METRIC.replace('ogbn-products-shape', f'ogbn-{args.shape}-shape')


def parse_args(argv=None):
  p = argparse.ArgumentParser()
  p.add_argument('--gpus', type=int, default=1)
  p.add_argument('--steps', type=int, default=50)
  p.add_argument('--warmup', type=int, default=5)
  p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  p.add_argument('--shape', default='products', choices=sorted(SHAPES))
  p.add_argument('--nodes', type=int, default=None)
  p.add_argument('--edges', type=int, default=None)
  p.add_argument('--feat-dim', type=int, default=None)
  p.add_argument('--classes', type=int, default=None)
  p.add_argument('--batch', type=int, default=1024)
  p.add_argument('--hidden', type=int, default=256)
  p.add_argument('--fanout', default='15,10,5')
  p.add_argument('--path', default='engine', choices=['engine', 'loader'],
                 help='ours: fused CUDA-graph engine, or NeighborLoader + eager model (the API the reference arm uses)')
  p.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'],
                 help='ours: compute dtype (the engine is bf16-only; fp32 runs the loader path)')
  p.add_argument('--ref-config', default='hbm', choices=['hbm', 'stock'],
                 help="reference arm: 'hbm' = graph_mode CUDA + split_ratio 1.0; 'stock' = ZERO_COPY + split_ratio 0.2")
  p.add_argument('--ref-dtype', default='bf16', choices=['bf16', 'fp32'])
  p.add_argument('--no-arms', action='store_true', help='skip the secondary pairings')
  p.add_argument('--min-time', type=float, default=1.0,
                 help='repeat the K-step timed region on fresh batches until this many seconds are measured')
  p.add_argument('--feat-format', default='bf16', choices=['bf16', 'mxfp8'],
                 help='ours/engine: feature storage (mxfp8 = e4m3 + UE8M0/32 block scales, de-quantised in the fused kernel)')
  p.add_argument('--dropout', type=float, default=0.0,
                 help='hidden-layer dropout in the engine (default 0: the reference arm model has none)')
  p.add_argument('--no-fused', action='store_true')
  p.add_argument('--fused', default='auto', choices=['auto', 'on'],
                 help="'auto': time fused vs unfused layer 1 at warm-up and keep the faster")
  p.add_argument('--no-graph', action='store_true')
  p.add_argument('--no-calibrate', action='store_true', help='size the arena for the worst case')
  p.add_argument('--no-pipeline', action='store_true', help='do not overlap sample(b+1) with train(b)')
  p.add_argument('--hot-fraction', type=float, default=None,
                 help='multi-GPU: fraction of every rank\'s (hotness-ordered) feature rows replicated on all GPUs '
                      '(default 0.25 for products, 0.15 for papers100m like the reference example)')
  p.add_argument('--replica-budget-gb', type=float, default=16.0,
                 help='multi-GPU placement policy: per-GPU HBM budget for REPLICATED data.  The CSR is replicated on '
                      'every GPU when it fits (the reference\'s multi-GPU layout: graph_mode=CUDA per trainer), the '
                      'rest of the budget replicates the hottest feature rows (NVSwitch multicast fill); what does '
                      'not fit stays range-partitioned and is read from peer HBM in-kernel.  0 = fully partitioned '
                      '(hot fraction 0.25 / 0.15 as in round 1)')
  p.add_argument('--seed', type=int, default=0)
  p.add_argument('--sections', action='store_true', help='print per-stage device times (eager) and exit')
  p.add_argument('--kernel-times', action='store_true',
                 help='print the in-situ device time of every launch of an eager step (warm L2, events) and exit')
  p.add_argument('--profile-steps', type=int, default=0,
                 help='run this many eager steps between cudaProfilerStart/Stop (for ncu) and exit')
  a = p.parse_args(argv)
  sh = SHAPES[a.shape]
  for k in ('nodes', 'edges', 'feat_dim', 'classes'):
    if getattr(a, k) is None:
      setattr(a, k, sh[k])
  a.replicate_topology = False
  a.placement = 'explicit --hot-fraction'
  if a.hot_fraction is None:
    budget = a.replica_budget_gb * 2 ** 30
    if budget <= 0:
      a.hot_fraction = 0.15 if a.shape == 'papers100m' else 0.25
      a.placement = 'partitioned (budget 0)'
    else:
      topo_bytes = a.edges * (4 if a.nodes < 2 ** 31 - 1 else 8)
      if topo_bytes <= budget:
        a.replicate_topology = True
        budget -= topo_bytes
      in_dim = (a.feat_dim + 63) // 64 * 64
      row_bytes = in_dim * 2 if a.feat_format == 'bf16' else in_dim + 16
      a.hot_fraction = min(1.0, budget / float(a.nodes * row_bytes))
      a.placement = f'replica budget {a.replica_budget_gb:g} GB/GPU'
  if a.dtype == 'fp32':
    a.path = 'loader'
  return a


def canonical_config(args, world):
  """The part of `config` that BOTH arms print identically: what is being measured."""
  return {
    'model': 'GraphSAGE-3x256-mean', 'global_batch': args.batch * world, 'seq_len': None,
    'fanout': args.fanout, 'hidden': args.hidden,
    'graph': f'RMAT({args.shape}-shape) nodes={args.nodes} directed_edges={(args.edges // 2) * 2}',
    'feat_dim': args.feat_dim, 'classes': args.classes, 'optimizer': 'Adam', 'loss': 'NLL',
    'parallelism': f'dp{world}',
    'memory_tier': 'hbm-resident graph+features',
    'precision': 'bf16 feature storage, fp32 neighbour-mean accumulation, bf16 tensor-core GEMMs, fp32 master weights',
    'timing': 'K-step region repeated on fresh seed batches until >= min_time s; inputs larger than L2',
  }


class ClockSampler(threading.Thread):
  """Polls SM clocks + throttle reasons during the timed region (NVML in-process at ~1 kHz;
  `nvidia-smi -lms` cannot resolve a region that lasts tens of milliseconds)."""

  def __init__(self, gpu_index=0):
    super().__init__(daemon=True)
    self.gpu_index = gpu_index
    self.samples = []          # (sm_mhz, reasons bitmask)
    self.sm_max = None
    self._stop_evt = threading.Event()
    self.gate = threading.Event()      # samples are kept only while set (the reference arm opens it around its timed
    self.gate.set()                    # regions; this arm samples from start() to stop())

  def run(self):
    try:
      import pynvml as nv
      nv.nvmlInit()
      h = nv.nvmlDeviceGetHandleByIndex(self.gpu_index)
      self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
      get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
          nv.nvmlDeviceGetCurrentClocksThrottleReasons
      while not self._stop_evt.is_set():
        if self.gate.is_set():
          self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(get_reasons(h))))
        time.sleep(0.001)
    except Exception:
      self._smi_fallback()

  def _smi_fallback(self):
    q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    while not self._stop_evt.is_set():
      if not self.gate.is_set():
        time.sleep(0.01)
        continue
      try:
        out = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i',
                              str(self.gpu_index)], capture_output=True, text=True, timeout=5).stdout
        p = [x.strip() for x in out.strip().split(',')]
        mask = 0
        for bit, val in zip((0x8, 0x40, 0x20, 0x4), p[2:6]):
          if val.lower().startswith('active'):
            mask |= bit
        self.samples.append((float(p[0]), mask))
        self.sm_max = float(p[1])
      except Exception:
        return

  def stop(self):
    self._stop_evt.set()

  def summary(self):
    # NVML reason bits: 0x4 sw_power_cap, 0x8 hw_slowdown, 0x20 sw_thermal, 0x40 hw_thermal
    names = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}
    if not self.samples:
      return {'sm_mhz': None, 'sm_max_mhz': self.sm_max, 'reasons': [], 'samples': 0}
    reasons = set()
    for _, m in self.samples:
      for bit, n in names.items():
        if m & bit:
          reasons.add(n)
    return {'sm_mhz': statistics.median(c for c, _ in self.samples), 'sm_max_mhz': self.sm_max,
            'reasons': sorted(reasons), 'samples': len(self.samples)}


def setup_dist(args):
  import torch
  import torch.distributed as dist
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local_rank)
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    # stdout carries exactly one JSON line: NCCL's banner / debug lines go to stderr
    os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
  return rank, world, local_rank


def max_over_ranks(x, world, device):
  import torch
  import torch.distributed as dist
  if world == 1:
    return x
  t = torch.tensor([x], dtype=torch.float64, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return float(t.item())


class BlockTimer(object):
  """Times `step(i)` in back-to-back blocks of K steps until `min_time` seconds of device time have been
  measured (CUDA events around every block, barrier + synchronize on both sides, max over ranks per block)."""

  def __init__(self, world, device, K, min_time, max_blocks=2000):
    self.world, self.device, self.K, self.min_time, self.max_blocks = world, device, K, min_time, max_blocks

  def barrier(self):
    import torch
    import torch.distributed as dist
    if self.world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def run(self, step):
    import torch
    blocks, i = [], 0
    while True:
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      self.barrier()
      e0.record()
      for _ in range(self.K):
        step(i)
        i += 1
      e1.record()
      self.barrier()
      blocks.append(max_over_ranks(e0.elapsed_time(e1), self.world, self.device))
      # every rank sees the same (max-reduced) block times, so all ranks stop together
      if sum(blocks) >= self.min_time * 1e3 or len(blocks) >= self.max_blocks:
        break
    total_ms = sum(blocks)
    return {'ms_per_step': total_ms / (len(blocks) * self.K), 'blocks': len(blocks),
            'steps_total': len(blocks) * self.K, 'seconds': total_ms / 1e3,
            'first_block_ms_per_step': blocks[0] / self.K,
            'median_block_ms_per_step': statistics.median(blocks) / self.K}


def build_ours(args, rank, world, device, need_engine=True):
  """Synthetic dataset of the requested shape -> (engine, train seed pool)."""
  import torch
  import graphlearn_for_pytorch_b200 as glt
  from graphlearn_for_pytorch_b200.models import GraphSageEngine
  from graphlearn_for_pytorch_b200.parallel import PartitionedFeature, PartitionedGraph, range_bounds
  from graphlearn_for_pytorch_b200.utils.synthetic import rmat_csr_shard, rmat_degrees

  N, E = args.nodes, args.edges
  in_dim = (args.feat_dim + 63) // 64 * 64        # 100 -> 128 (zero-padded columns)
  half = E // 2
  bounds = range_bounds(N, world)
  old2new = None
  if world > 1 and 0 < args.hot_fraction < 1.0:
    # hotness reordering (the reference's examples do sort_by_in_degree + split_ratio): nodes are
    # sorted by degree and dealt round-robin to the ranks, so every rank's id range is balanced and
    # starts with its hottest rows, which are replicated on all GPUs through NVSwitch multicast
    from graphlearn_for_pytorch_b200.parallel import hotness_balanced_order
    deg = rmat_degrees(N, half, seed=args.seed, device=device, undirected=True)
    old2new, bounds = hotness_balanced_order(deg, world)
    del deg
  # every rank streams the same RMAT edge sequence (same seed) and keeps its own row range only: the
  # full edge list (26 GB at papers100M shape) is never materialised
  shard = rmat_csr_shard(N, half, bounds[rank], bounds[rank + 1], seed=args.seed, device=device,
                         undirected=True, old2new=old2new)
  del old2new
  torch.cuda.empty_cache()
  g = torch.Generator(device=device)
  g.manual_seed(args.seed + 1)
  labels = torch.randint(0, args.classes, (N,), device=device, generator=g)
  if world == 1:
    graph = glt.data.Graph.from_shards([shard], device.index)
    feats = torch.zeros(N, in_dim, dtype=torch.bfloat16, device=device)
    rows = 1 << 22
    for b0 in range(0, N, rows):      # chunked: no fp32 copy of the whole table
      b1 = min(N, b0 + rows)
      feats[b0:b1, :args.feat_dim] = torch.randn(b1 - b0, args.feat_dim, device=device, generator=g).to(torch.bfloat16)
    if args.feat_format == 'mxfp8':
      feats = torch.cat([glt.data.quantize_mxfp8(feats[b0:min(N, b0 + rows)]) for b0 in range(0, N, rows)])
    ut = glt.data.UnifiedTensor(device.index, feats.dtype)
    ut.append_shared_tensor(feats)
    table = ut._table()
    keep = (graph, ut, feats, shard)
  else:
    pg = PartitionedGraph(shard, bounds, device, replicate_topology=args.replicate_topology)
    graph = pg.graph
    b, e = bounds[rank], bounds[rank + 1]
    gl = torch.Generator(device=device)
    gl.manual_seed(args.seed + 100 + rank)
    local = torch.zeros(e - b, in_dim, dtype=torch.bfloat16, device=device)
    rows = 1 << 22
    for b0 in range(0, e - b, rows):
      b1 = min(e - b, b0 + rows)
      local[b0:b1, :args.feat_dim] = torch.randn(b1 - b0, args.feat_dim, device=device, generator=gl).to(torch.bfloat16)
    if args.feat_format == 'mxfp8':
      local = torch.cat([glt.data.quantize_mxfp8(local[b0:min(e - b, b0 + rows)]) for b0 in range(0, e - b, rows)])
    pf = PartitionedFeature(local, bounds, device,
                            hot_per_rank=min(int(args.hot_fraction * (bounds[1] - bounds[0])),
                                             min(bounds[r + 1] - bounds[r] for r in range(world))),
                            full_replica=args.hot_fraction >= 1.0)
    table = pf.table
    keep = (pg, pf)
  torch.cuda.empty_cache()
  fanouts = [int(x) for x in args.fanout.split(',')]
  # training seeds: each rank draws from its own slice of a fixed permutation (DDP-style)
  gp = torch.Generator(device='cpu')
  gp.manual_seed(args.seed + 7)
  perm = torch.randperm(N, generator=gp)
  pool = perm[rank::world]
  eng = GraphSageEngine(graph, table, labels, in_dim=in_dim, num_nodes=N, fanouts=fanouts,
                        batch_size=args.batch, hidden=args.hidden, num_classes=args.classes,
                        lr=3e-3, seed=args.seed, device=device,
                        use_fused=False if args.no_fused else (True if args.fused == 'on' else 'auto'),
                        use_cuda_graph=not args.no_graph,
                        calibration_seeds=None if args.no_calibrate else pool,
                        pipeline=not args.no_pipeline, feature_format=args.feat_format, dropout=args.dropout)
  eng._keep = keep
  return eng, pool


def seed_batches(pool, n_batches, bs):
  """[n_batches, bs] fresh seed batches: consecutive slices of the rank's permutation, wrapping into
  further (re-shuffled) epochs when the run is longer than one epoch."""
  import torch
  need = n_batches * bs
  parts, have, ep = [], 0, 0
  while have < need:
    if ep == 0:
      p = pool
    else:
      g = torch.Generator(device='cpu')
      g.manual_seed(1000 + ep)
      p = pool[torch.randperm(pool.numel(), generator=g)]
    parts.append(p)
    have += p.numel()
    ep += 1
  return torch.cat(parts)[:need].view(n_batches, bs).contiguous()


def run_loader_arm(args, device, dtype, min_time, K, W):
  """This framework through the SAME public API the reference arm uses: Dataset + NeighborLoader(as_pyg_v1)
  + the reference arm's plain-PyTorch SAGE model code (baseline/ref_bench.py::_sage_model) + torch Adam."""
  import torch
  import torch.nn.functional as F
  import graphlearn_for_pytorch_b200 as glt
  from graphlearn_for_pytorch_b200.utils.synthetic import rmat_edges
  sys.path.insert(0, os.path.join(ROOT, 'baseline'))
  from ref_bench import _sage_model
  N, E = args.nodes, args.edges
  # host tensors handed to the Dataset API, exactly like the reference arm does
  ei = rmat_edges(N, E // 2, seed=args.seed, device=device)
  ei = torch.cat([ei, ei.flip(0)], 1).cpu()
  g = torch.Generator()
  g.manual_seed(args.seed + 1)
  labels = torch.randint(0, args.classes, (N,), generator=g)
  fdt = torch.bfloat16 if dtype == 'bf16' else torch.float32
  feats = torch.randn(N, args.feat_dim, generator=g).to(fdt)
  ds = glt.data.Dataset()
  ds.init_graph(ei, graph_mode='CUDA', directed=False, device=device.index)
  del ei
  ds.init_node_features(feats, sort_func=glt.data.sort_by_in_degree, split_ratio=1.0, with_gpu=True,
                        device=device.index, dtype=fdt)
  ds.init_node_labels(labels)
  labels = labels.to(device)
  gp = torch.Generator(device='cpu')
  gp.manual_seed(args.seed + 7)
  pool = torch.randperm(N, generator=gp)
  bs = args.batch
  fan = [int(x) for x in args.fanout.split(',')]
  model = _sage_model(torch, args.feat_dim, args.hidden, args.classes).to(device)
  opt = torch.optim.Adam(model.parameters(), lr=3e-3)
  lab = labels
  loss_host = torch.zeros(1).pin_memory()
  state = {'it': None}

  def new_iter():
    g2 = torch.Generator(device='cpu')
    g2.manual_seed(int(time.time_ns() % (1 << 31)))
    seeds = pool[torch.randperm(pool.numel(), generator=g2)]
    loader = glt.loader.NeighborLoader(ds, fan, seeds, batch_size=bs, shuffle=False, drop_last=True,
                                       device=device, as_pyg_v1=True)
    state['it'] = iter(loader)

  def step(read_loss):
    try:
      batch_size, n_id, adjs = next(state['it'])
    except (StopIteration, TypeError):
      new_iter()
      batch_size, n_id, adjs = next(state['it'])
    x = ds.node_features[n_id]
    opt.zero_grad(set_to_none=True)
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=(dtype == 'bf16')):
      out = model(x, adjs)
    loss = F.nll_loss(out.float(), lab[n_id[:batch_size]])
    loss.backward()
    opt.step()
    if read_loss:
      loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
    return loss

  new_iter()
  for _ in range(W):
    step(False)
  bt = BlockTimer(1, device, K, min_time)
  dev_t = bt.run(lambda i: step(False))
  e2e_t = bt.run(lambda i: step(True))
  del ds, feats, model, opt
  torch.cuda.empty_cache()
  return {'path': 'loader', 'api': 'Dataset + loader.NeighborLoader(as_pyg_v1=True) + plain-PyTorch SAGE + torch Adam',
          'dtype': dtype, 'value': bs / (dev_t['ms_per_step'] / 1e3), 'ms_per_step': dev_t['ms_per_step'],
          'e2e_value': bs / (e2e_t['ms_per_step'] / 1e3), 'e2e_ms_per_step': e2e_t['ms_per_step'],
          'timed': dev_t, 'pairs_with': f'--impl reference --ref-config hbm --ref-dtype {dtype}'}


def run_ours(args):
  import torch
  rank, world, local_rank = setup_dist(args)
  device = torch.device('cuda', local_rank)
  import torch.distributed as dist
  bs, K, W = args.batch, args.steps, args.warmup

  if args.path == 'loader':
    # the loader path is a single-process measurement (each rank would hold a full replica)
    arm = run_loader_arm(args, device, args.dtype, args.min_time, K, W) if rank == 0 else None
    if rank == 0:
      cfg = canonical_config(args, 1)
      cfg['precision'] = 'fp32' if args.dtype == 'fp32' else cfg['precision']
      print(json.dumps({
        'metric': metric_name(args), 'value': arm['value'], 'unit': 'samples/s', 'n_gpus': 1, 'steps': K, 'warmup': W,
        'ms_per_step': arm['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': arm['value'] / BASELINE_SAMPLES_PER_S, 'dtype': args.dtype, 'data': 'synthetic',
        'impl': 'ours', 'config': cfg, 'details': {'path': 'loader', 'api': arm['api']}, 'timed': arm['timed'],
        'e2e': {'value': arm['e2e_value'], 'unit': 'samples/s', 'ms_per_step': arm['e2e_ms_per_step'],
                'h2d_bytes_per_step': bs * 8, 'd2h_bytes_per_step': 4},
        'gpu_launches': None}), flush=True)
    return

  eng, pool = build_ours(args, rank, world, device)
  eng.warmup_and_capture(n_eager=2)
  if args.sections:
    eng._graphs = []
    nb = max(1, min(12, pool.numel() // bs))
    sec = eng.profile_sections(pool[:nb * bs].view(nb, bs).to(device), iters=10)
    t = torch.tensor([sec[k] for k in sorted(sec)], device=device)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
      print(json.dumps({'sections_ms_max_over_ranks': dict(zip(sorted(sec), [round(float(x), 4) for x in t])),
                        'n_gpus': world}), flush=True)
    if world > 1:
      dist.barrier()
      os._exit(0)
    return
  if args.kernel_times:
    eng._graphs = []
    nb = max(1, min(12, pool.numel() // bs))
    kt = eng.profile_kernels(pool[:nb * bs].view(nb, bs).to(device), iters=10)
    if rank == 0:
      print('# in-situ device time per launch (eager unpipelined step, events, mean of 10 steps)')
      for name, us in kt:
        print(f'{us:9.2f} us  {name}')
      print(f'{sum(u for _, u in kt):9.2f} us  total')
    if world > 1:
      dist.barrier()
      os._exit(0)
    return
  if args.profile_steps > 0:
    # ncu --profile-from-start off: only these eager steps are captured
    eng._graph_fb = eng._graph_opt = eng._graph_full = None
    eng._graphs = []
    sd = seed_batches(pool, 3 + args.profile_steps, bs).to(device)
    for i in range(3):
      eng.train_step(sd[i])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for i in range(args.profile_steps):
      eng.train_step(sd[3 + i])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    return

  # enough fresh batches for the longest plausible run: min_time at ~0.15 ms/step, x2 safety; wraps beyond
  n_pool_batches = max(K + W, min(int(args.min_time / 0.15e-3) + K, 40000))
  seeds_all = seed_batches(pool, n_pool_batches, bs)
  seeds_dev = seeds_all.to(device)
  seeds_pinned = seeds_all.pin_memory()
  loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
  bt = BlockTimer(world, device, K, args.min_time)

  # ---------------- kernel-only (device-resident seeds) ----------------
  for i in range(W):
    eng.train_step(seeds_dev[i % n_pool_batches])
  bt.barrier()
  clocks = ClockSampler(local_rank)
  if rank == 0:
    clocks.start()
  dev_t = bt.run(lambda i: eng.train_step(seeds_dev[(W + i) % n_pool_batches]))
  last_loss = float(eng.loss.item())

  # ---------------- end-to-end through the public API ----------------
  for i in range(W):
    eng.train_step(seeds_pinned[i % n_pool_batches])

  def e2e_step(i):
    loss = eng.train_step(seeds_pinned[(W + i) % n_pool_batches])   # H2D of this step's seeds inside
    loss_host.copy_(loss, non_blocking=True)                         # D2H read of this step's loss
  e2e_t = bt.run(e2e_step)
  if rank == 0:
    clocks.stop()
  overflow = eng.overflow_count()
  c = eng.arena.counters.cpu().tolist()

  arms = []
  if rank == 0 and world == 1 and not args.no_arms and args.shape == 'products':
    # secondary pairings through the loader API (the engine's buffers stay allocated: products shape is small)
    for dt in ('bf16', 'fp32'):
      try:
        arms.append(run_loader_arm(args, device, dt, min(args.min_time, 0.5), K, W))
      except Exception as ex:  # never lose the headline because a secondary arm failed
        arms.append({'path': 'loader', 'dtype': dt, 'error': f'{type(ex).__name__}: {str(ex)[:200]}'})

  if rank == 0:
    per_step_seeds = bs * world
    value = per_step_seeds / (dev_t['ms_per_step'] / 1e3)
    e2e = per_step_seeds / (e2e_t['ms_per_step'] / 1e3)
    out = {
      'metric': metric_name(args), 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': K, 'warmup': W,
      'ms_per_step': dev_t['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': value / BASELINE_SAMPLES_PER_S, 'dtype': 'bf16', 'data': 'synthetic',
      'impl': 'ours',
      'config': canonical_config(args, world),
      'details': {
        'path': 'engine (GraphSageEngine.train_step)', 'feat_dim_padded': eng.in_dim,
        'partitioning': f'graph/feature range-partition over {world} GPU(s), in-kernel P2P' if world > 1 else 'single GPU',
        'fused_tcgen05_layer1': bool(eng.fused_ok[1]),
        'layer1_autotune_ms': getattr(eng, 'autotune_ms', {}).get(1),
        'tcgen05_gemm_layers': getattr(eng, 'tc_gemm', None),
        'gather_backward': bool(eng.use_gather_bwd),
        'feature_format': args.feat_format,
        'remote_rows_staged_on_sampling_stream': bool(getattr(eng, 'stage_remote', False)),
        'hot_feature_replica': (None if world == 1 else {'fraction': round(args.hot_fraction, 4),
                                                         'fill': getattr(eng._keep[1], 'fill_mode', None)}),
        'placement': (None if world == 1 else {'policy': args.placement,
                                               'topology_replicated': bool(args.replicate_topology)}),
        'grad_allreduce': 'peer-HBM all-reduce fused into Adam (NVLink, in-graph)' if eng.peer_group is not None
                          else ('nccl' if world > 1 else 'none'),
        'cuda_graph': eng._graph_fb is not None, 'pipelined_sample_train_overlap': bool(eng.pipeline),
        'l2_policy': 'inputs larger than L2 (feature table + CSR >> 126 MB, fresh random seed batch every step)',
        'baseline_ref': 'BASELINE.md GraphSAGE papers100M epoch 8.56 s / 1,207,179 seeds on 4xA100 (derived; context only)',
        'last_batch_nodes': c[1:5], 'last_batch_edges': c[6:9], 'last_loss': last_loss,
        'arena': {'calibrated': bool(getattr(eng, 'calibrated', False)), 'cap_rows': [int(x) for x in eng.cap_rows],
                  'dropped_neighbours_total': int(overflow)},
      },
      'timed': dev_t,
      'e2e': {'value': e2e, 'unit': 'samples/s', 'ms_per_step': e2e_t['ms_per_step'],
              'h2d_bytes_per_step': bs * 8, 'd2h_bytes_per_step': 4, 'timed': e2e_t,
              'api': 'GraphSageEngine.train_step(pinned host seeds) + loss D2H'},
      'gpu_launches': int(eng.kernels_per_step) * K,
      'kernels_per_step': int(eng.kernels_per_step),
      'library_gemm_launches_per_step': int(getattr(eng, 'library_gemms_per_step', -1)),
      'arms': arms,
      'clocks': clocks.summary(),
    }
    print(json.dumps(out), flush=True)
  if world > 1:
    # orderly teardown, but never let a stuck NCCL/IPC teardown hold the box: every rank has
    # reported, so after a final barrier a hard exit is safe
    eng.close()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    watchdog = threading.Timer(20.0, lambda: os._exit(0))
    watchdog.daemon = True
    watchdog.start()
    dist.destroy_process_group()
    watchdog.cancel()


def run_reference(args):
  ref_dir = os.path.join(ROOT, 'baseline', '_ref')
  if not os.path.isdir(os.path.join(ref_dir, 'graphlearn_torch')):
    print(json.dumps({'impl': 'reference', 'unavailable': 'baseline/_ref not installed'}))
    return
  if args.shape != 'products':
    print(json.dumps({'impl': 'reference', 'unavailable':
                      'the reference builds its CSR and feature store on the host from a full COO edge list; at '
                      'papers100M shape (1.6 B edges, 57 GB fp32 features) that exceeds the bench time budget'}))
    return
  try:
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import ref_bench
    try:       # SM clocks / throttle reasons of the reference arm's timed regions (same sampler as this arm's)
      clocks = ClockSampler(int(os.environ.get('LOCAL_RANK', '0')))
      clocks.gate.clear()
      clocks.start()
      args._clock_sampler = clocks
    except Exception:
      args._clock_sampler = None
    ref_bench.main(args, BASELINE_SAMPLES_PER_S, canonical_config(args, int(os.environ.get('WORLD_SIZE', '1'))), metric_name(args))
  except Exception as e:  # the reference arm must never break the driver
    if int(os.environ.get('RANK', '0')) == 0:
      print(json.dumps({'impl': 'reference', 'unavailable': f'{type(e).__name__}: {str(e)[:200]}'}))


if __name__ == '__main__':
  a = parse_args()
  if a.impl == 'reference':
    run_reference(a)
  else:
    run_ours(a)
