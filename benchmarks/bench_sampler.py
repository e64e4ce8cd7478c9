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
p.add_argument('--impl', default='ours', choices=['ours', 'reference'],
               help="'reference': the unmodified reference (baseline/_ref) NeighborSampler on the same box and graph")
args = p.parse_args()

if args.impl == 'reference':
  # reference benchmarks/api/bench_sampler.py:27-54, same synthetic graph, same seeds, device-timed
  ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  sys.path.insert(0, os.path.join(ROOT, 'baseline', 'shims'))
  sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))
  try:
    import graphlearn_torch as rglt
    dev = torch.device('cuda', 0)
    ei = rmat_edges(args.nodes, args.edges // 2, seed=0, device=dev)
    ei = torch.cat([ei, ei.flip(0)], 1).cpu()
    csr = rglt.data.Topology(ei, input_layout='COO')
    g = rglt.data.Graph(csr, args.mode, 0)
    fan = [int(v) for v in args.fanout.split(',')]
    sampler = rglt.sampler.NeighborSampler(g, fan, device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(0)
    seeds = [torch.randint(0, args.nodes, (args.batch,), device=dev, generator=gen) for _ in range(args.iters + 5)]
    for sd in seeds[:5]:
      sampler.sample_from_nodes(sd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    edges = 0
    e0.record()
    for sd in seeds[5:]:
      edges += sampler.sample_from_nodes(sd).row.numel()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({'impl': 'reference', 'metric': 'sampled edges/s', 'graph_mode': args.mode, 'batch': args.batch,
                      'fanout': fan, 'edges_per_batch': edges / args.iters, 'api_M_edges_per_s': edges / ms / 1e3,
                      'api_ms_per_batch': ms / args.iters}))
  except Exception as ex:  # noqa: BLE001
    print(json.dumps({'impl': 'reference', 'unavailable': f'{type(ex).__name__}: {str(ex)[:300]}'}))
  sys.exit(0)

dev = torch.device('cuda', 0)
ei = rmat_edges(args.nodes, args.edges // 2, seed=0, device=dev)
ei = torch.cat([ei, ei.flip(0)], 1)
topo = glt.data.Topology(ei if args.mode == 'CUDA' else ei.cpu(), layout='CSR', num_nodes=args.nodes)
del ei
graph = glt.data.Graph(topo, args.mode, 0)
fan = [int(v) for v in args.fanout.split(',')]
sampler = NeighborSampler(graph, fan, device=dev, seed=1)
gen = torch.Generator(device=dev); gen.manual_seed(0)
seeds = [torch.randint(0, args.nodes, (args.batch,), device=dev, generator=gen) for _ in range(args.iters + 5)]
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
print(json.dumps({'impl': 'ours', 'metric': 'sampled edges/s', 'graph_mode': args.mode, 'batch': args.batch, 'fanout': fan,
                  'edges_per_batch': edges / args.iters,
                  'api_M_edges_per_s': edges / api_ms / 1e3, 'api_ms_per_batch': api_ms / args.iters,
                  'arena_M_edges_per_s': edges / arena_ms / 1e3, 'arena_ms_per_batch': arena_ms / args.iters,
                  'reference_published_A100_M_edges_per_s': 32.08}))
