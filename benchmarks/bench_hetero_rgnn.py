"""BASELINE config 4: R-GNN on an IGBH-shaped heterogeneous graph with per-edge-type (cross-partition)
sampling.  1 GPU: plain Graphs.  N GPUs (torchrun): every relation is range-partitioned over the GPUs
and sampled through peer-HBM reads; features are partitioned per node type.

  python benchmarks/bench_hetero_rgnn.py --papers 1000000
  torchrun --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/bench_hetero_rgnn.py --papers 4000000
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'examples'))
import graphlearn_for_pytorch_b200 as glt  # noqa: E402
from common import synthetic_igbh  # noqa: E402
from graphlearn_for_pytorch_b200.models import RGNN  # noqa: E402
from graphlearn_for_pytorch_b200.parallel import PartitionedFeature, partition_hetero_graph  # noqa: E402
from graphlearn_for_pytorch_b200.sampler import NeighborSampler, NodeSamplerInput  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument('--papers', type=int, default=200_000)
p.add_argument('--batch', type=int, default=1024)
p.add_argument('--fanout', default='15,10,5')
p.add_argument('--steps', type=int, default=30)
p.add_argument('--warmup', type=int, default=3)
p.add_argument('--model', default='rsage')
args = p.parse_args()
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
  dist.init_process_group('nccl', device_id=dev)
edges, feats, labels, sizes = synthetic_igbh(args.papers, args.papers // 2, max(args.papers // 400, 8),
                                              max(args.papers // 1000, 8), feat_dim=128)
edge_dir = 'in'
topos = {et: glt.data.Topology(ei.to(dev), layout='CSC', num_nodes=sizes[et[2]]) for et, ei in edges.items()}
if world > 1:
  graphs, bounds, keep = partition_hetero_graph(topos, sizes, rank, world, dev, edge_dir)
  fstore = {nt: PartitionedFeature(feats[nt][bounds[nt][rank]:bounds[nt][rank + 1]].to(dev).to(torch.bfloat16),
                                   bounds[nt], dev) for nt in sizes}
else:
  graphs = {et: glt.data.Graph(t, 'CUDA', local) for et, t in topos.items()}
  fstore = {}
  for nt in sizes:
    ut = glt.data.UnifiedTensor(local, torch.bfloat16); ut.append_shared_tensor(feats[nt].to(dev).to(torch.bfloat16))
    fstore[nt] = ut
fan = [int(v) for v in args.fanout.split(',')]
sampler = NeighborSampler(graphs, fan, device=dev, edge_dir=edge_dir, seed=1)
y = labels['paper'].to(dev)
out0 = sampler.sample_from_nodes(NodeSamplerInput(torch.arange(args.batch, device=dev), 'paper'))
model = RGNN(list(out0.row.keys()), 128, 256, int(y.max()) + 1, num_layers=len(fan), node_type='paper',
             model=args.model).to(dev).to(torch.bfloat16)
if world > 1:
  model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
pool = torch.randperm(args.papers, generator=torch.Generator().manual_seed(3))[rank::world].to(dev)


def step(i):
  seeds = pool[(i * args.batch) % (pool.numel() - args.batch):][:args.batch]
  out = sampler.sample_from_nodes(NodeSamplerInput(seeds, 'paper'))
  x = {nt: fstore[nt][ids] for nt, ids in out.node.items()}
  ei = {et: torch.stack([out.row[et], out.col[et]]) for et in out.row}
  logits = model(x, ei, out.num_sampled_nodes, out.num_sampled_edges)[:out.batch['paper'].numel()].float()
  loss = F.cross_entropy(logits, y[out.batch['paper']])
  opt.zero_grad(); loss.backward(); opt.step()
  return loss, sum(v.numel() for v in out.node.values()), sum(v.numel() for v in out.row.values())


for i in range(args.warmup):
  step(i)
torch.cuda.synchronize()
if world > 1:
  dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
nodes = edges_n = 0
for i in range(args.steps):
  loss, n, e = step(args.warmup + i)
  nodes += n; edges_n += e
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
  dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
  print(json.dumps({'metric': 'R-GNN igbh-shape training throughput (seed nodes/s, device-timed, max over ranks)',
                    'value': args.steps * args.batch * world / (ms.item() / 1e3), 'n_gpus': world,
                    'ms_per_step': ms.item() / args.steps, 'model': args.model, 'papers': args.papers,
                    'fanout': fan, 'nodes_per_batch': nodes / args.steps, 'edges_per_batch': edges_n / args.steps,
                    'loss': float(loss.detach()), 'path': 'hetero NeighborSampler (per-edge-type device sampling, P2P shards) '
                    '+ eager RGNN (bf16)'}))
if world > 1:
  dist.barrier()
  dist.destroy_process_group()
