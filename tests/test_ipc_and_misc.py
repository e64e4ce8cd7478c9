import os

import torch
import torch.multiprocessing as mp

import graphlearn_for_pytorch_b200 as glt
from graphlearn_for_pytorch_b200.sampler import NeighborSampler
from dist_utils import run_workers
from helpers import ring_dataset


def _child(ds, q):
  # the Dataset travelled through ForkingPickler reducers (shared-memory tensors)
  s = NeighborSampler(ds.graph, [2, 2], seed=1)
  out = s.sample_from_nodes(torch.tensor([0, 10]))
  x = ds.node_features[out.node]
  q.put((out.node.tolist(), x[:, 0].tolist(), ds.node_labels[out.node].tolist()))


def test_dataset_ipc_to_spawned_process():
  ds = ring_dataset(40)
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  p = ctx.Process(target=_child, args=(ds, q))
  p.start()
  nodes, feat0, labels = q.get(timeout=120)
  p.join(60)
  assert p.exitcode == 0
  assert feat0 == [float(n) for n in nodes] and labels == nodes and nodes[:2] == [0, 10]


def test_random_seed_manager_reproducible():
  from graphlearn_for_pytorch_b200.utils import RandomSeedManager
  _, topo = __import__('helpers').rmat_csr(500, 8000)
  g = glt.data.Graph(topo, 'CPU')
  outs = []
  for _ in range(2):
    RandomSeedManager.set_seed(42)
    s = NeighborSampler(g, [3, 2])
    outs.append(s.sample_from_nodes(torch.arange(50)).node)
  assert torch.equal(outs[0], outs[1])
  RandomSeedManager._seed = None


def test_weighted_neighbor_sampler_cpu():
  ds = ring_dataset(40, weights=True)
  s = NeighborSampler(ds.graph, [1], with_weight=True, seed=3)
  hits = {}
  for t in range(600):
    o = s.sample_one_hop(torch.tensor([0]), 1)
    hits[int(o.nbr)] = hits.get(int(o.nbr), 0) + 1
  # node 0 -> 1 (edge 0, weight 1) and -> 2 (edge 1, weight 2)
  assert set(hits) == {1, 2} and 1.4 < hits[2] / hits[1] < 2.9


def _w_all2all(rank, world, port):
  import torch.distributed as dist
  import graphlearn_for_pytorch_b200.distributed as d
  from dist_utils import build_partition
  dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
  ds = build_partition(rank, world)
  df = d.DistFeature(world, rank, ds.node_features, ds.node_feat_pb, local_only=True)
  df.local_only = False
  ids = torch.tensor([3, 0, 7, 39, 38, 1, 1, 20 + rank])
  out = df.get_all2all(ids)
  assert torch.equal(out[:, 0].long(), ids), out[:, 0]
  dist.barrier()
  dist.destroy_process_group()


def test_dist_feature_all2all_gloo():
  run_workers(_w_all2all)
