"""Offline partitioning of a dataset into the on-disk layout consumed by DistDataset.load
(counterpart of the reference's examples/distributed/partition_ogbn_dataset.py: hotness from
NeighborSampler.sample_prob -> FrequencyPartitioner with a per-partition hot-feature cache)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import add_dataset_args, glt, load_homo  # noqa: E402
from graphlearn_for_pytorch_b200.partition import FrequencyPartitioner, RandomPartitioner  # noqa: E402
from graphlearn_for_pytorch_b200.sampler import NeighborSampler, NodeSamplerInput  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument('--out', required=True)
p.add_argument('--parts', type=int, default=2)
add_dataset_args(p, nodes=50_000, edges=500_000)   # --root <dir with ogbn_products/>: partition the real dataset
p.add_argument('--strategy', default='frequency', choices=['frequency', 'random'])
p.add_argument('--cache-ratio', type=float, default=0.05)
args = p.parse_args()

ei, x, y, split, args.nodes = load_homo(args)
os.makedirs(args.out, exist_ok=True)
torch.save(y, os.path.join(args.out, 'labels.pt'))
train = split['train']
torch.save(train, os.path.join(args.out, 'train_idx.pt'))
for name in ('valid', 'test'):                      # evaluation seeds travel with the partitions, as in the reference
  if name in split:                                 # (examples/distributed/partition_ogbn_dataset.py:62-84)
    torch.save(split[name], os.path.join(args.out, f'{name}_idx.pt'))
if args.strategy == 'random':
  RandomPartitioner(args.out, args.parts, args.nodes, ei, node_feat=x).partition()
else:
  topo = glt.data.Topology(ei, layout='CSR', num_nodes=args.nodes)
  graph = glt.data.Graph(topo, 'CUDA' if torch.cuda.is_available() else 'CPU')
  sampler = NeighborSampler(graph, [15, 10, 5])
  probs = [sampler.sample_prob(NodeSamplerInput(train[r::args.parts]), args.nodes).cpu() for r in range(args.parts)]
  FrequencyPartitioner(args.out, args.parts, args.nodes, ei, probs, node_feat=x, cache_ratio=args.cache_ratio,
                       chunk_size=20_000).partition()
print('partitioned into', args.out)
