"""BASELINE config 5: SEAL-style enclosing-subgraph sampling + strict negative sampling throughput on a
large RMAT graph (range-partitioned over the GPUs when launched with torchrun).

Metrics: links/s for (2-seed, k-hop `-1`-fanout capped) induced subgraph extraction, and M negatives/s
for strict negative sampling -- both reading peer shards in-kernel in multi-GPU runs (the reference
falls back to non-strict local negatives and RPC-broadcast node sets in distributed mode).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphlearn_for_pytorch_b200 as glt  # noqa: E402
from graphlearn_for_pytorch_b200.parallel import PartitionedGraph, range_bounds  # noqa: E402
from graphlearn_for_pytorch_b200.sampler import NeighborSampler, NodeSamplerInput, RandomNegativeSampler  # noqa: E402
from graphlearn_for_pytorch_b200.utils.synthetic import rmat_csr_shard  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument('--nodes', type=int, default=10_000_000)
p.add_argument('--edges', type=int, default=100_000_000)
p.add_argument('--links', type=int, default=256, help='links per batch (one joint subgraph per batch)')
p.add_argument('--fanout', default='20,20', help='neighbour cap per hop when growing the enclosing subgraph')
p.add_argument('--iters', type=int, default=20)
args = p.parse_args()
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
  dist.init_process_group('nccl', device_id=dev)
# every rank streams the same RMAT edge sequence and keeps its own row range only (1 B directed edges never exist as
# one edge list); `--edges` counts DIRECTED edges of the symmetrised graph like the other benchmarks
bounds = range_bounds(args.nodes, world)
shard = rmat_csr_shard(args.nodes, args.edges // 2, bounds[rank], bounds[rank + 1], seed=0, device=dev, undirected=True)
torch.cuda.empty_cache()
if world > 1:
  pg = PartitionedGraph(shard, bounds, dev)
  graph = pg.graph
else:
  graph = glt.data.Graph.from_shards([shard], local)
graph._col_count = args.nodes
local_edges = int(shard['indices'].numel())
fan = [int(v) for v in args.fanout.split(',')]
sampler = NeighborSampler(graph, fan, device=dev, with_edge=False, seed=1)
neg = RandomNegativeSampler(graph, 'CUDA', seed=2)
neg.graph._col_count = args.nodes


def timed(fn, iters):
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  tot = 0
  for _ in range(iters):
    tot += fn()
  e1.record(); torch.cuda.synchronize()
  ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
  if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  return ms.item(), tot


def sub():
  seeds = torch.randint(0, args.nodes, (2 * args.links,), device=dev)
  out = sampler.subgraph(NodeSamplerInput(seeds))
  return out.row.numel()


def negs():
  return neg.sample(1 << 20, trials_num=5).shape[1]


ms_s, sub_edges = timed(sub, args.iters)
ms_n, n_neg = timed(negs, args.iters)
if rank == 0:
  print(json.dumps({'metric': 'SEAL subgraph + negative sampling', 'n_gpus': world, 'nodes': args.nodes,
                    'edges': args.edges, 'edges_on_rank0': local_edges, 'links_per_s': args.iters * args.links * world / (ms_s / 1e3),
                    'induced_edges_per_batch': sub_edges / args.iters, 'ms_per_subgraph_batch': ms_s / args.iters,
                    'strict_negatives_M_per_s': n_neg * world / ms_n / 1e3}))
if world > 1:
  dist.barrier()
  dist.destroy_process_group()
