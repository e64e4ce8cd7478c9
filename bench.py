#!/usr/bin/env python
"""Headline benchmark: GraphSAGE training throughput (seed nodes / s) on an
ogbn-products-shape synthetic graph (BASELINE.json metric/config).

  python bench.py --gpus N --steps K --warmup W            # this framework
  python bench.py --impl reference --gpus N ...             # unmodified reference (baseline/_ref)

Model/config (reference examples/train_sage_ogbn_products.py:30-59,112-114): 3-layer GraphSAGE,
hidden 256, fanout [15,10,5], batch 1024 seeds per GPU, Adam, NLL loss; graph = RMAT with the
ogbn-products shape (2,449,029 nodes, 123.7 M directed edges, 100-dim features, 47 classes),
random-init weights, synthetic features/labels (no network access for the real dataset).

Prints ONE JSON line on rank 0.  `value` is device-timed (CUDA events, max over ranks) over
exactly K steps with seeds already resident on the device; `e2e` re-measures the same metric
through the public `GraphSageEngine.train_step(pinned_host_seeds)` call including the per-step
H2D seed copy and the per-step D2H loss read.
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
# throughput of the reference).
BASELINE_SAMPLES_PER_S = 1_207_179 / 8.56


def parse_args():
  p = argparse.ArgumentParser()
  p.add_argument('--gpus', type=int, default=1)
  p.add_argument('--steps', type=int, default=50)
  p.add_argument('--warmup', type=int, default=5)
  p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  p.add_argument('--nodes', type=int, default=2_449_029)
  p.add_argument('--edges', type=int, default=123_718_280)
  p.add_argument('--feat-dim', type=int, default=100)
  p.add_argument('--classes', type=int, default=47)
  p.add_argument('--batch', type=int, default=1024)
  p.add_argument('--hidden', type=int, default=256)
  p.add_argument('--fanout', default='15,10,5')
  p.add_argument('--no-fused', action='store_true')
  p.add_argument('--fused', default='auto', choices=['auto', 'on'],
                 help="'auto': time fused vs unfused layer 1 at warm-up and keep the faster")
  p.add_argument('--no-graph', action='store_true')
  p.add_argument('--no-calibrate', action='store_true', help='size the arena for the worst case')
  p.add_argument('--no-pipeline', action='store_true', help='do not overlap sample(b+1) with train(b)')
  p.add_argument('--hot-fraction', type=float, default=0.25,
                 help='multi-GPU: fraction of every rank\'s (hotness-ordered) feature rows replicated on all GPUs')
  p.add_argument('--seed', type=int, default=0)
  p.add_argument('--sections', action='store_true', help='print per-stage device times (eager) and exit')
  p.add_argument('--profile-steps', type=int, default=0,
                 help='run this many eager steps between cudaProfilerStart/Stop (for ncu) and exit')
  return p.parse_args()


class ClockSampler(threading.Thread):
  """Polls SM clocks + throttle reasons during the timed region (NVML in-process at ~1 kHz;
  `nvidia-smi -lms` cannot resolve a region that lasts tens of milliseconds)."""

  def __init__(self, gpu_index=0):
    super().__init__(daemon=True)
    self.gpu_index = gpu_index
    self.samples = []          # (sm_mhz, reasons bitmask)
    self.sm_max = None
    self._stop_evt = threading.Event()

  def run(self):
    try:
      import pynvml as nv
      nv.nvmlInit()
      h = nv.nvmlDeviceGetHandleByIndex(self.gpu_index)
      self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
      get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
          nv.nvmlDeviceGetCurrentClocksThrottleReasons
      while not self._stop_evt.is_set():
        self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(get_reasons(h))))
        time.sleep(0.001)
    except Exception:
      self._smi_fallback()

  def _smi_fallback(self):
    q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    while not self._stop_evt.is_set():
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


def build_ours(args, rank, world, device):
  """Synthetic products-shape dataset -> (engine, train seed pool)."""
  import torch
  import graphlearn_for_pytorch_b200 as glt
  from graphlearn_for_pytorch_b200.models import GraphSageEngine
  from graphlearn_for_pytorch_b200.parallel import (PartitionedFeature, PartitionedGraph, range_bounds,
                                                     shard_topology)
  from graphlearn_for_pytorch_b200.utils.synthetic import rmat_edges

  N, E = args.nodes, args.edges
  in_dim = (args.feat_dim + 63) // 64 * 64        # 100 -> 128 (zero-padded columns)
  # every rank generates the same graph (same seed), then keeps its row range only
  half = E // 2
  ei = rmat_edges(N, half, seed=args.seed, device=device)
  ei = torch.cat([ei, ei.flip(0)], dim=1)         # undirected, like ogbn-products
  bounds = range_bounds(N, world)
  if world > 1 and args.hot_fraction > 0:
    # hotness reordering (the reference's example does sort_by_in_degree + split_ratio): nodes are
    # sorted by degree and dealt round-robin to the ranks, so every rank's id range is balanced and
    # starts with its hottest rows, which are replicated on all GPUs through NVSwitch multicast
    from graphlearn_for_pytorch_b200.parallel import hotness_balanced_order
    deg = torch.bincount(ei[0], minlength=N)
    old2new, bounds = hotness_balanced_order(deg, world)
    ei = old2new[ei]
    del deg, old2new
  topo = glt.data.Topology(ei, layout='CSR', num_nodes=N)
  del ei
  g = torch.Generator(device=device)
  g.manual_seed(args.seed + 1)
  labels = torch.randint(0, args.classes, (N,), device=device, generator=g)
  if world == 1:
    graph = glt.data.Graph(topo, 'CUDA', device.index)
    graph.lazy_init()
    feats = torch.zeros(N, in_dim, dtype=torch.bfloat16, device=device)
    feats[:, :args.feat_dim] = torch.randn(N, args.feat_dim, device=device, generator=g).to(torch.bfloat16)
    ut = glt.data.UnifiedTensor(device.index, torch.bfloat16)
    ut.append_shared_tensor(feats)
    table = ut._table()
    keep = (graph, ut, feats)
  else:
    shard = shard_topology(topo, bounds, rank, device)
    shard['eids'] = None
    pg = PartitionedGraph(shard, bounds, device)
    graph = pg.graph
    b, e = bounds[rank], bounds[rank + 1]
    # identical full-feature RNG stream on every rank would cost N x F; generate per-shard
    gl = torch.Generator(device=device)
    gl.manual_seed(args.seed + 100 + rank)
    local = torch.zeros(e - b, in_dim, dtype=torch.bfloat16, device=device)
    local[:, :args.feat_dim] = torch.randn(e - b, args.feat_dim, device=device, generator=gl).to(torch.bfloat16)
    pf = PartitionedFeature(local, bounds, device,
                            hot_per_rank=int(args.hot_fraction * (bounds[1] - bounds[0])))
    table = pf.table
    keep = (pg, pf)
  del topo
  torch.cuda.empty_cache()
  fanouts = [int(x) for x in args.fanout.split(',')]
  # training seeds: each rank draws from its own slice of a fixed permutation (DDP-style)
  gp = torch.Generator(device='cpu')
  gp.manual_seed(args.seed + 7)
  perm = torch.randperm(N, generator=gp)
  pool = perm[rank::world]
  eng = GraphSageEngine(graph, table, labels, in_dim=in_dim, num_nodes=N, fanouts=fanouts,
                        batch_size=args.batch, hidden=args.hidden, num_classes=args.classes,
                        lr=3e-3, seed=args.seed, device=device, use_fused=False if args.no_fused else (True if args.fused == 'on' else 'auto'),
                        use_cuda_graph=not args.no_graph,
                        calibration_seeds=None if args.no_calibrate else pool,
                        pipeline=not args.no_pipeline)
  eng._keep = keep
  return eng, pool


def run_ours(args):
  import torch
  rank, world, local_rank = setup_dist(args)
  device = torch.device('cuda', local_rank)
  import torch.distributed as dist
  eng, pool = build_ours(args, rank, world, device)
  bs, K, W = args.batch, args.steps, args.warmup
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
  if args.profile_steps > 0:
    # ncu --profile-from-start off: only these eager steps are captured
    eng._graph_fb = eng._graph_opt = eng._graph_full = None
    eng._graphs = []
    sd = pool[:bs].to(device)
    for _ in range(3):
      eng.train_step(sd)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(args.profile_steps):
      eng.train_step(sd)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    return

  n_batches = K + W
  need = n_batches * bs
  reps = (need + pool.numel() - 1) // pool.numel()
  seeds_all = pool.repeat(reps)[:need].view(n_batches, bs).contiguous()
  seeds_dev = seeds_all.to(device)
  seeds_pinned = seeds_all.pin_memory()
  loss_host = torch.zeros(K + W, dtype=torch.float32).pin_memory()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # ---------------- kernel-only (device-resident seeds) ----------------
  for i in range(W):
    eng.train_step(seeds_dev[i])
  barrier()
  clocks = ClockSampler(local_rank)
  if rank == 0:
    clocks.start()
  t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  t0.record()
  for i in range(K):
    eng.train_step(seeds_dev[W + i])
  t1.record()
  barrier()
  ms = max_over_ranks(t0.elapsed_time(t1), world, device)
  last_loss = float(eng.loss.item())

  # ---------------- end-to-end through the public API ----------------
  for i in range(W):
    eng.train_step(seeds_pinned[i])
  barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  e0.record()
  for i in range(K):
    loss = eng.train_step(seeds_pinned[W + i])            # H2D of this step's seeds inside
    loss_host[i:i + 1].copy_(loss, non_blocking=True)      # D2H read of this step's loss
  e1.record()
  barrier()
  e2e_ms = max_over_ranks(e0.elapsed_time(e1), world, device)
  if rank == 0:
    clocks.stop()

  if rank == 0:
    total = K * bs * world
    value = total / (ms / 1e3)
    e2e = total / (e2e_ms / 1e3)
    c = eng.arena.counters.cpu().tolist()
    out = {
      'metric': 'GraphSAGE ogbn-products-shape training throughput (seed nodes/s, device-timed, max over ranks)',
      'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': K, 'warmup': W,
      'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': value / BASELINE_SAMPLES_PER_S, 'dtype': 'bf16', 'data': 'synthetic',
      'impl': 'ours',
      'config': {
        'model': 'GraphSAGE-3x256-mean', 'global_batch': bs * world, 'seq_len': None,
        'fanout': args.fanout, 'graph': f'RMAT nodes={args.nodes} directed_edges={(args.edges // 2) * 2}',
        'feat_dim': args.feat_dim, 'feat_dim_padded': eng.in_dim, 'classes': args.classes,
        'parallelism': f'dp{world}+graph/feature range-partition over {world} GPU(s), in-kernel P2P',
        'optimizer': 'Adam(fused)', 'fused_tcgen05_layer1': bool(eng.fused_ok[1]),
        'layer1_autotune_ms': getattr(eng, 'autotune_ms', {}).get(1),
        'hot_feature_replica': (None if world == 1 else {'fraction': args.hot_fraction, 'fill': getattr(eng._keep[1], 'fill_mode', None)}),
        'grad_allreduce': 'peer-HBM all-reduce fused into Adam (NVLink, in-graph)' if eng.peer_group is not None else ('nccl' if world > 1 else 'none'),
        'cuda_graph': eng._graph_fb is not None, 'pipelined_sample_train_overlap': bool(eng.pipeline),
        'l2_policy': 'inputs larger than L2 (feature table + CSR >> 126 MB, random rows per batch)',
        'baseline_ref': 'BASELINE.md GraphSAGE papers100M epoch 8.56 s / 1,207,179 seeds on 4xA100 (derived)',
        'last_batch_nodes': c[1:5], 'last_batch_edges': c[6:9], 'last_loss': last_loss,
        'arena': {'calibrated': bool(getattr(eng, 'calibrated', False)), 'cap_rows': [int(x) for x in eng.cap_rows],
                  'dropped_neighbours_total': int(c[12])},
      },
      'e2e': {'value': e2e, 'unit': 'samples/s', 'ms_per_step': e2e_ms / K,
              'h2d_bytes_per_step': bs * 8, 'd2h_bytes_per_step': 4},
      'gpu_launches': int(eng.kernels_per_step) * K,
      'kernels_per_step': int(eng.kernels_per_step),
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
  try:
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import ref_bench
    ref_bench.main(args, BASELINE_SAMPLES_PER_S)
  except Exception as e:  # the reference arm must never break the driver
    if int(os.environ.get('RANK', '0')) == 0:
      print(json.dumps({'impl': 'reference', 'unavailable': f'{type(e).__name__}: {str(e)[:200]}'}))


if __name__ == '__main__':
  a = parse_args()
  if a.impl == 'reference':
    run_reference(a)
  else:
    run_ours(a)
