"""Distributed loader throughput (nodes / edges / feature-rows per second per epoch), the metric
of the reference's benchmarks/api/bench_dist_neighbor_loader.py:139-162.  Run one process per
partition (see examples/distributed/dist_train_sage.py for the launch pattern)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphlearn_for_pytorch_b200.distributed as gd  # noqa: E402


def main():
  p = argparse.ArgumentParser()
  p.add_argument('--root', required=True)
  p.add_argument('--rank', type=int, required=True)
  p.add_argument('--world', type=int, default=2)
  p.add_argument('--master-addr', default='127.0.0.1')
  p.add_argument('--master-port', type=int, default=29900)
  p.add_argument('--workers', type=int, default=2)
  p.add_argument('--epochs', type=int, default=2)
  args = p.parse_args()
  os.environ.setdefault('MASTER_ADDR', args.master_addr)
  os.environ.setdefault('MASTER_PORT', str(args.master_port))
  cuda = torch.cuda.is_available()
  device = torch.device('cuda', args.rank % max(torch.cuda.device_count(), 1)) if cuda else torch.device('cpu')
  gd.init_worker_group(args.world, args.rank)
  ds = gd.DistDataset().load(args.root, args.rank, graph_mode='CUDA' if cuda else 'CPU', feature_with_gpu=cuda,
                             device=device.index)
  train = torch.load(os.path.join(args.root, 'train_idx.pt'))
  train = train[ds.node_pb[train] == args.rank]
  opts = gd.MpDistSamplingWorkerOptions(num_workers=args.workers, worker_concurrency=4, master_addr=args.master_addr,
                                        master_port=args.master_port + 1, pin_memory=cuda)
  loader = gd.DistNeighborLoader(ds, [15, 10, 5], train, batch_size=1024, shuffle=True, collect_features=True,
                                 to_device=device, worker_options=opts)
  for epoch in range(args.epochs):
    t0, nodes, edges = time.time(), 0, 0
    for b in loader:
      nodes += b.node.numel(); edges += b.edge_index.shape[1]
    dt = time.time() - t0
    print(f'[rank {args.rank}] epoch {epoch}: {nodes / dt / 1e6:.3f} M nodes(+features)/s, {edges / dt / 1e6:.3f} M edges/s')
  loader.shutdown()


if __name__ == '__main__':   # spawned sampling workers re-import this module
  main()
