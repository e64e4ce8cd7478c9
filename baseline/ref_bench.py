"""Reference arm of bench.py: the UNMODIFIED reference (installed from /root/reference into
baseline/_ref) run through its own public API and stock code path (glt.data.Dataset + Feature +
glt.loader.NeighborLoader), following its examples/train_sage_ogbn_products.py:95-141 (3-layer
GraphSAGE hidden 256, NeighborLoader([15,10,5], batch 1024, as_pyg_v1=True), Adam, NLL).  Same
synthetic products-shape graph as our arm.

Pairings (bench.py --ref-config / --ref-dtype):
  hbm   : graph_mode='CUDA', split_ratio=1.0  -- the reference's own fully HBM-resident recipe
          (examples/train_sage_prod_with_trim.py:97,103); default, pairs with our HBM-resident arm
  stock : graph_mode='ZERO_COPY', split_ratio=0.2 -- the products example as shipped
  bf16  : bf16 feature storage (the reference's UnifiedTensor registers BFloat16,
          csrc/cuda/unified_tensor.cu:129) + torch.autocast(bf16) around the model, fp32 master weights
  fp32  : the example's stock precision

PyG is not installable offline, so (a) `torch_sparse` / `torch_geometric.data` are satisfied by
the tiny shims in baseline/shims (dependencies only), and (b) the SAGEConv layers of the example
are written in plain PyTorch below (same math as PyG's SAGEConv(mean)).  None of this repo's
kernels, engine or package is imported here.
"""
import json
import os
import statistics
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def _sage_model(torch, in_dim, hidden, out_dim):
  nn, F = torch.nn, torch.nn.functional

  class SAGEConv(nn.Module):
    def __init__(self, i, o):
      super().__init__()
      self.lin_l = nn.Linear(i, o)
      self.lin_r = nn.Linear(i, o, bias=False)

    def forward(self, x, edge_index):
      x_src, x_dst = x
      src, dst = edge_index[0], edge_index[1]
      # the neighbour mean is accumulated in fp32 whatever the storage dtype (bf16 index_add_ atomics are
      # ~2x slower than fp32 ones on this GPU and lose precision; measured: 8.5 ms vs 4.5 ms per step), the two
      # Linear layers then run in the autocast dtype -- the fastest mixed-precision recipe for this model
      xs = x_src.float()
      agg = torch.zeros(x_dst.shape[0], xs.shape[1], dtype=torch.float32, device=xs.device)
      agg.index_add_(0, dst, xs[src])
      deg = torch.zeros(x_dst.shape[0], dtype=torch.float32, device=xs.device)
      deg.index_add_(0, dst, torch.ones_like(dst, dtype=torch.float32))
      mean = (agg / deg.clamp(min=1).unsqueeze(1)).to(self.lin_l.weight.dtype)   # no-op unless the model is pure bf16
      return self.lin_l(mean) + self.lin_r(x_dst.to(self.lin_r.weight.dtype) if not torch.is_autocast_enabled() else x_dst)

  class SAGE(nn.Module):
    def __init__(self):
      super().__init__()
      self.convs = nn.ModuleList([SAGEConv(in_dim, hidden), SAGEConv(hidden, hidden), SAGEConv(hidden, out_dim)])

    def forward(self, x, adjs):
      for i, (edge_index, _, size) in enumerate(adjs):
        x_target = x[:size[1]]
        x = self.convs[i]((x, x_target), edge_index)
        if i != len(self.convs) - 1:
          x = F.relu(x)
      return x.float().log_softmax(dim=-1)

  return SAGE()


def main(args, baseline_samples_per_s, canonical_config, metric):
  sys.path.insert(0, os.path.join(HERE, 'shims'))
  sys.path.insert(0, os.path.join(HERE, '_ref'))
  import torch
  import torch.distributed as dist
  import torch.nn.functional as F
  import graphlearn_torch as glt  # the unmodified reference

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=device)

  N, E = args.nodes, args.edges
  # same generator as our arm (pure torch, duplicated here to keep this arm independent)
  def rmat(num_nodes, num_edges, seed, a=0.57, b=0.19, c=0.19):
    scale = max(1, (num_nodes - 1).bit_length())
    gen = torch.Generator(device=device); gen.manual_seed(seed)
    outs, done, chunk = [], 0, 1 << 24
    while done < num_edges:
      n = min(chunk, num_edges - done)
      src = torch.zeros(n, dtype=torch.int64, device=device); dst = torch.zeros_like(src)
      for _ in range(scale):
        r = torch.rand(n, device=device, generator=gen)
        src = (src << 1) | (r >= a + b).to(torch.int64)
        dst = (dst << 1) | ((r >= a) & (r < a + b) | (r >= a + b + c)).to(torch.int64)
      outs.append(torch.stack([src % num_nodes, dst % num_nodes])); done += n
    ei = torch.cat(outs, 1)
    return (ei * 0x9E3779B1 + 12345) % num_nodes
  ei = rmat(N, E // 2, args.seed)
  ei = torch.cat([ei, ei.flip(0)], 1).cpu()
  g = torch.Generator(); g.manual_seed(args.seed + 1)
  labels = torch.randint(0, args.classes, (N,), generator=g)
  feats = torch.randn(N, args.feat_dim, generator=g)

  cfg_name, dt_name = args.ref_config, args.ref_dtype
  graph_mode, split = ('CUDA', 1.0) if cfg_name == 'hbm' else ('ZERO_COPY', 0.2)
  ds = glt.data.Dataset()
  ds.init_graph(edge_index=ei, graph_mode=graph_mode, directed=False, device=local_rank)
  del ei
  ds.init_node_labels(node_label_data=labels)
  node_labels = ds.node_labels.to(device)
  gp = torch.Generator(); gp.manual_seed(args.seed + 7)
  pool = torch.randperm(N, generator=gp)[rank::world]
  bs, K, W = args.batch, args.steps, args.warmup
  fan = [int(x) for x in args.fanout.split(',')]

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def max_ranks(ms):
    if world > 1:
      t = torch.tensor([ms], dtype=torch.float64, device=device)
      dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    return ms

  def measure(dtype_name, min_time, pure=False):
    """Builds Feature + loader + model for one precision and times it (same block protocol as our arm).
    bf16 comes in two recipes: torch.autocast around an fp32 model (casts per call), or `pure`: the model itself
    in bf16 (no per-call weight casts; torch Adam on bf16 parameters) -- the faster one is the headline."""
    fdt = torch.bfloat16 if dtype_name == 'bf16' else torch.float32
    ds.node_features = None
    ds.init_node_features(node_feature_data=feats.to(fdt), sort_func=glt.data.sort_by_in_degree, split_ratio=split,
                          device_group_list=[glt.data.DeviceGroup(0, [local_rank])], device=local_rank)
    torch.manual_seed(args.seed)
    model = _sage_model(torch, args.feat_dim, args.hidden, args.classes).to(device)
    if pure:
      model = model.to(torch.bfloat16)
    if world > 1:
      model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])
    opt = torch.optim.Adam(model.parameters(), lr=3e-3)
    loss_host = torch.zeros(1).pin_memory()
    state = {'it': None, 'ep': 0}

    def new_iter():
      g2 = torch.Generator(); g2.manual_seed(1000 + state['ep']); state['ep'] += 1
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
      adjs = [adj.to(device) for adj in adjs]
      opt.zero_grad()
      with torch.autocast('cuda', dtype=torch.bfloat16, enabled=(dtype_name == 'bf16' and not pure)):
        out = model(ds.node_features[n_id], adjs)
      loss = F.nll_loss(out, node_labels[n_id[:batch_size]])
      loss.backward()
      opt.step()
      if read_loss:
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)

    def timed(read_loss):
      blocks = []
      sampler = getattr(args, '_clock_sampler', None)
      if sampler is not None:
        sampler.gate.set()               # SM-clock samples are taken inside the timed regions only
      while True:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); e0.record()
        for _ in range(K):
          step(read_loss)
        e1.record(); barrier()
        blocks.append(max_ranks(e0.elapsed_time(e1)))
        if sum(blocks) >= min_time * 1e3 or len(blocks) >= 2000:
          break
      if sampler is not None:
        sampler.gate.clear()
      tot = sum(blocks)
      return {'ms_per_step': tot / (len(blocks) * K), 'blocks': len(blocks), 'steps_total': len(blocks) * K,
              'seconds': tot / 1e3, 'first_block_ms_per_step': blocks[0] / K,
              'median_block_ms_per_step': statistics.median(blocks) / K}

    new_iter()
    for _ in range(W):
      step(False)
    d, e = timed(False), timed(True)
    del model, opt
    return d, e

  arms = []
  recipe = 'fp32'
  if dt_name == 'bf16':
    # two bf16 recipes; the FASTER one is the headline of this arm, the other is listed under "arms"
    cand = {}
    for name, pure in (('bf16-autocast', False), ('bf16-pure', True)):
      try:
        cand[name] = measure('bf16', args.min_time, pure=pure)
      except Exception as ex:
        arms.append({'ref_config': cfg_name, 'dtype': name, 'error': f'{type(ex).__name__}: {str(ex)[:200]}'})
    recipe = min(cand, key=lambda k: cand[k][0]['ms_per_step'])
    dev_t, e2e_t = cand[recipe]
    for name, (d2, e2) in cand.items():
      if name != recipe:
        arms.append({'ref_config': cfg_name, 'dtype': name, 'value': bs * world / (d2['ms_per_step'] / 1e3),
                     'ms_per_step': d2['ms_per_step'], 'e2e_value': bs * world / (e2['ms_per_step'] / 1e3),
                     'e2e_ms_per_step': e2['ms_per_step'], 'timed': d2})
  else:
    dev_t, e2e_t = measure(dt_name, args.min_time)
  if not args.no_arms and dt_name == 'bf16':
    try:
      d2, e2 = measure('fp32', min(args.min_time, 0.5))
      arms.append({'ref_config': cfg_name, 'dtype': 'fp32', 'value': bs * world / (d2['ms_per_step'] / 1e3),
                   'ms_per_step': d2['ms_per_step'], 'e2e_value': bs * world / (e2['ms_per_step'] / 1e3),
                   'e2e_ms_per_step': e2['ms_per_step'], 'timed': d2})
    except Exception as ex:
      arms.append({'ref_config': cfg_name, 'dtype': 'fp32', 'error': f'{type(ex).__name__}: {str(ex)[:200]}'})
  if rank == 0:
    per_step = bs * world
    val = per_step / (dev_t['ms_per_step'] / 1e3)
    cfg = dict(canonical_config)
    if cfg_name != 'hbm':
      cfg['memory_tier'] = 'ZERO_COPY topology (pinned host, UVA) + 20 % of feature rows in HBM'
    if dt_name != 'bf16':
      cfg['precision'] = 'fp32'
    print(json.dumps({
      'impl': 'reference', 'metric': metric,
      'value': val, 'unit': 'samples/s', 'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': dev_t['ms_per_step'],
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': val / baseline_samples_per_s,
      'dtype': 'bf16' if dt_name == 'bf16' else 'fp32', 'data': 'synthetic',
      'config': cfg,
      'details': {'model_code': f'plain-PyTorch SAGEConv (PyG unavailable offline), recipe {recipe} (fastest bf16 recipe of '
                                'this run: autocast around an fp32 model vs the model itself in bf16)' if dt_name == 'bf16'
                  else 'plain-PyTorch SAGEConv (PyG unavailable offline), fp32',
                  'graph_mode': graph_mode, 'feature_split_ratio': split, 'feature_dtype': dt_name,
                  'loader': 'glt.loader.NeighborLoader(as_pyg_v1=True)', 'parallelism': f'ddp{world}',
                  'deps': 'torch_sparse/torch_geometric.data shims (baseline/shims)'},
      'timed': dev_t,
      'e2e': {'value': per_step / (e2e_t['ms_per_step'] / 1e3), 'unit': 'samples/s', 'ms_per_step': e2e_t['ms_per_step'],
              'h2d_bytes_per_step': bs * 8, 'd2h_bytes_per_step': 4, 'timed': e2e_t},
      'arms': arms,
      'clocks': args._clock_sampler.summary() if getattr(args, '_clock_sampler', None) is not None else None,
    }), flush=True)
  if world > 1:
    dist.destroy_process_group()
