"""Distributed GraphSAGE, worker mode: every rank loads one partition, samples across partitions
(RPC data plane, usable across machines) and trains with DDP.

Counterpart of the reference's examples/distributed/dist_train_sage_supervised.py.
  python examples/distributed/partition_dataset.py --out /tmp/parts --parts 2
  for r in 0 1: python examples/distributed/dist_train_sage.py --root /tmp/parts --rank r --world 2 &
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import glt  # noqa: E402
import graphlearn_for_pytorch_b200.distributed as gd  # noqa: E402
from graphlearn_for_pytorch_b200.models import GraphSAGE  # noqa: E402


def main():
  p = argparse.ArgumentParser()
  p.add_argument('--root', required=True)
  p.add_argument('--rank', type=int, required=True)
  p.add_argument('--world', type=int, default=2)
  p.add_argument('--master-addr', default='127.0.0.1')
  p.add_argument('--master-port', type=int, default=29700)
  p.add_argument('--epochs', type=int, default=2)
  p.add_argument('--workers', type=int, default=0, help='0 = collocated sampling, >0 = sampling subprocesses')
  args = p.parse_args()

  cuda = torch.cuda.is_available()
  device = torch.device('cuda', args.rank % max(torch.cuda.device_count(), 1)) if cuda else torch.device('cpu')
  os.environ.setdefault('MASTER_ADDR', args.master_addr)
  os.environ.setdefault('MASTER_PORT', str(args.master_port))
  dist.init_process_group('nccl' if cuda else 'gloo', rank=args.rank, world_size=args.world)
  gd.init_worker_group(args.world, args.rank)
  ds = gd.DistDataset()
  ds.load(args.root, args.rank, graph_mode='CUDA' if cuda else 'CPU', feature_with_gpu=cuda,
          whole_node_label_file=os.path.join(args.root, 'labels.pt'), device=device.index)
  train = torch.load(os.path.join(args.root, 'train_idx.pt'))
  train = train[ds.node_pb[train] == args.rank]
  if args.workers > 0:
    opts = gd.MpDistSamplingWorkerOptions(num_workers=args.workers, worker_concurrency=4, master_addr=args.master_addr,
                                          master_port=args.master_port + 1, pin_memory=cuda)
  else:
    opts = gd.CollocatedDistSamplingWorkerOptions(master_addr=args.master_addr, master_port=args.master_port + 1)
  loader = gd.DistNeighborLoader(ds, [15, 10, 5], train, batch_size=512, shuffle=True, drop_last=False,
                                 collect_features=True, to_device=device, worker_options=opts)
  n_cls = int(ds.node_labels.max()) + 1
  model = torch.nn.parallel.DistributedDataParallel(GraphSAGE(ds.node_features.shape[1], 256, n_cls, 3).to(device))
  opt = torch.optim.Adam(model.parameters(), lr=3e-3)
  for epoch in range(args.epochs):
    for b in loader:
      out = model(b.x, b.edge_index, b.num_sampled_nodes, b.num_sampled_edges)[:b.batch_size]
      loss = F.cross_entropy(out, b.y[:b.batch_size])
      opt.zero_grad(); loss.backward(); opt.step()
    print(f'[rank {args.rank}] epoch {epoch} loss {float(loss.detach()):.4f}')
  loader.shutdown()
  dist.barrier()
  if gd.rpc_is_initialized():
    gd.barrier()
    gd.shutdown_rpc()


if __name__ == '__main__':   # spawned sampling workers re-import this module
  main()
