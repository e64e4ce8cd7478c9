"""Reference arm of bench.py: the UNMODIFIED reference (installed from /root/reference into
baseline/_ref) run through its own public API and stock code path, following its
examples/train_sage_ogbn_products.py:95-141 (Dataset.init_graph ZERO_COPY, Feature with
sort_by_in_degree + split_ratio 0.2, NeighborLoader([15,10,5], batch 1024, as_pyg_v1=True),
3-layer GraphSAGE hidden 256, Adam, NLL).  Same synthetic products-shape graph as our arm.

PyG is not installable offline, so (a) `torch_sparse` / `torch_geometric.data` are satisfied by
the tiny shims in baseline/shims (dependencies only), and (b) the SAGEConv layers of the example
are written in plain PyTorch below (same math as PyG's SAGEConv(mean)).  None of this repo's
kernels, engine or package is imported here.
"""
import json
import os
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
      agg = torch.zeros(x_dst.shape[0], x_src.shape[1], dtype=x_src.dtype, device=x_src.device)
      agg.index_add_(0, dst, x_src[src])
      deg = torch.zeros(x_dst.shape[0], dtype=x_src.dtype, device=x_src.device)
      deg.index_add_(0, dst, torch.ones_like(dst, dtype=x_src.dtype))
      return self.lin_l(agg / deg.clamp(min=1).unsqueeze(1)) + self.lin_r(x_dst)

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
      return x.log_softmax(dim=-1)

  return SAGE()


def main(args, baseline_samples_per_s):
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

  ds = glt.data.Dataset()
  ds.init_graph(edge_index=ei, graph_mode='ZERO_COPY', directed=False, device=local_rank)
  ds.init_node_features(node_feature_data=feats, sort_func=glt.data.sort_by_in_degree, split_ratio=0.2,
                        device_group_list=[glt.data.DeviceGroup(0, [local_rank])], device=local_rank)
  ds.init_node_labels(node_label_data=labels)
  gp = torch.Generator(); gp.manual_seed(args.seed + 7)
  pool = torch.randperm(N, generator=gp)[rank::world]
  bs, K, W = args.batch, args.steps, args.warmup
  need = (K + W) * bs * 2
  reps = (need + pool.numel() - 1) // pool.numel()
  seeds = pool.repeat(reps)[:need]
  fan = [int(x) for x in args.fanout.split(',')]
  loader = glt.loader.NeighborLoader(ds, fan, seeds, batch_size=bs, shuffle=False, drop_last=True,
                                     device=device, as_pyg_v1=True)
  model = _sage_model(torch, args.feat_dim, args.hidden, args.classes).to(device)
  if world > 1:
    model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])
  opt = torch.optim.Adam(model.parameters(), lr=3e-3)
  ds.node_labels = ds.node_labels.to(device)
  it = iter(loader)
  loss_host = torch.zeros(1).pin_memory()

  def step():
    batch_size, n_id, adjs = next(it)
    adjs = [adj.to(device) for adj in adjs]
    opt.zero_grad()
    out = model(ds.node_features[n_id], adjs)
    loss = F.nll_loss(out, ds.node_labels[n_id[:batch_size]])
    loss.backward()
    opt.step()
    return loss

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(n, read_loss):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); e0.record()
    for _ in range(n):
      loss = step()
      if read_loss:
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
    e1.record(); barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
      t = torch.tensor([ms], dtype=torch.float64, device=device)
      dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    return ms

  for _ in range(W):
    step()
  ms = timed(K, False)
  e2e_ms = timed(K, True)
  if rank == 0:
    total = K * bs * world
    val = total / (ms / 1e3)
    print(json.dumps({
      'impl': 'reference', 'metric': 'GraphSAGE ogbn-products-shape training throughput (seed nodes/s, device-timed, max over ranks)',
      'value': val, 'unit': 'samples/s', 'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms / K,
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': val / baseline_samples_per_s,
      'dtype': 'fp32 (reference stock path)', 'data': 'synthetic',
      'config': {'model': 'GraphSAGE-3x256-mean (plain PyTorch SAGEConv; PyG unavailable offline)',
                 'global_batch': bs * world, 'fanout': args.fanout, 'graph_mode': 'ZERO_COPY',
                 'feature_split_ratio': 0.2, 'loader': 'glt.loader.NeighborLoader(as_pyg_v1=True)',
                 'parallelism': f'ddp{world}', 'deps': 'torch_sparse/torch_geometric.data shims (baseline/shims)'},
      'e2e': {'value': total / (e2e_ms / 1e3), 'unit': 'samples/s', 'ms_per_step': e2e_ms / K,
              'h2d_bytes_per_step': bs * 8, 'd2h_bytes_per_step': 4},
    }), flush=True)
  if world > 1:
    dist.destroy_process_group()
