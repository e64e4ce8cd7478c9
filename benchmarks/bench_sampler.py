"""Neighbour-sampling throughput (M sampled edges / s), device-timed.

Counterpart of the reference's benchmarks/api/bench_sampler.py:27-54 (products, batch 1024,
fanout [15,10,5]; the reference reports 32.08 M edges/s on one A100 for papers100M).
Two numbers: `api` = NeighborSampler.sample_from_nodes (PyG-shaped output, one size read-back per
batch), `arena` = the static-shape arena alone (what the fused training engine uses; no host sync).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphlearn_for_pytorch_b200 as glt  # noqa: E402
from graphlearn_for_pytorch_b200.sampler import NeighborSampler  # noqa: E402
from graphlearn_for_pytorch_b200.utils.synthetic import rmat_edges  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument('--nodes', type=int, default=2_449_029)
p.add_argument('--edges', type=int, default=123_718_280)
p.add_argument('--batch', type=int, default=1024)
p.add_argument('--fanout', default='15,10,5')
p.add_argument('--iters', type=int, default=100)
p.add_argument('--mode', default='CUDA', choices=['CUDA', 'ZERO_COPY'])
args = p.parse_args()

dev = torch.device('cuda', 0)
ei = rmat_edges(args.nodes, args.edges // 2, seed=0, device=dev)
ei = torch.cat([ei, ei.flip(0)], 1)
topo = glt.data.Topology(ei if args.mode == 'CUDA' else ei.cpu(), layout='CSR', num_nodes=args.nodes)
del ei
graph = glt.data.Graph(topo, args.mode, 0)
fan = [int(v) for v in args.fanout.split(',')]
sampler = NeighborSampler(graph, fan, device=dev, seed=1)
seeds = [torch.randint(0, args.nodes, (args.batch,), device=dev) for _ in range(args.iters + 5)]
for s in seeds[:5]:
  sampler.sample_from_nodes(s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
edges = 0
e0.record()
for s in seeds[5:]:
  edges += sampler.sample_from_nodes(s).row.numel()
e1.record(); torch.cuda.synchronize()
api_ms = e0.elapsed_time(e1)
arena = sampler._arena
h = graph.graph_handler
e0.record()
for i, s in enumerate(seeds[5:]):
  arena.sample(h, s, None, 1, i * 8, False, False, False)
e1.record(); torch.cuda.synchronize()
arena_ms = e0.elapsed_time(e1)
print(json.dumps({'metric': 'sampled edges/s', 'graph_mode': args.mode, 'batch': args.batch, 'fanout': fan,
                  'edges_per_batch': edges / args.iters,
                  'api_M_edges_per_s': edges / api_ms / 1e3, 'api_ms_per_batch': api_ms / args.iters,
                  'arena_M_edges_per_s': edges / arena_ms / 1e3, 'arena_ms_per_batch': arena_ms / args.iters,
                  'reference_published_A100_M_edges_per_s': 32.08}))
