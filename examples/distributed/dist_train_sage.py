"""Distributed GraphSAGE, worker mode: every rank loads one partition, samples across partitions
(RPC data plane, usable across machines) and trains with DDP.

Counterpart of the reference's examples/distributed/dist_train_sage_supervised.py.
  python examples/distributed/partition_dataset.py --out /tmp/parts --parts 2
  for r in 0 1: python examples/distributed/dist_train_sage.py --root /tmp/parts --rank r --world 2 &
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch.distributed.algorithms.join import Join

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
  p.add_argument('--eval-every', type=int, default=1, help='epochs between test-set evaluations (0 = never)')
  p.add_argument('--batch', type=int, default=512)
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

  def worker_options(port):    # every loader has its own sampling-worker group, hence its own rendezvous port
    if args.workers > 0:
      return gd.MpDistSamplingWorkerOptions(num_workers=args.workers, worker_concurrency=4,
                                            master_addr=args.master_addr, master_port=port, pin_memory=cuda)
    return gd.CollocatedDistSamplingWorkerOptions(master_addr=args.master_addr, master_port=port)

  loader = gd.DistNeighborLoader(ds, [15, 10, 5], train, batch_size=args.batch, shuffle=True, drop_last=False,
                                 collect_features=True, to_device=device,
                                 worker_options=worker_options(args.master_port + 1))
  test_file = os.path.join(args.root, 'test_idx.pt')
  test_loader = None
  if args.eval_every > 0 and os.path.exists(test_file):
    test = torch.load(test_file)
    test = test[ds.node_pb[test] == args.rank]         # each rank scores the test seeds of its own partition
    test_loader = gd.DistNeighborLoader(ds, [15, 10, 5], test, batch_size=args.batch, shuffle=False, drop_last=False,
                                        collect_features=True, to_device=device,
                                        worker_options=worker_options(args.master_port + 2))
  n_cls = int(ds.node_labels.max()) + 1
  model = torch.nn.parallel.DistributedDataParallel(GraphSAGE(ds.node_features.shape[1], 256, n_cls, 3).to(device))
  opt = torch.optim.Adam(model.parameters(), lr=3e-3)

  @torch.no_grad()
  def test_accuracy():
    # reference examples/distributed/dist_train_sage_supervised.py:31-52: per-rank hits summed over the group.
    # The un-wrapped module is used: ranks see different numbers of test batches, DDP forward hooks must not fire.
    net = model.module
    net.eval()
    stat = torch.zeros(2, device=device)
    for b in test_loader:
      out = net(b.x, b.edge_index, b.num_sampled_nodes, b.num_sampled_edges)[:b.batch_size]
      stat[0] += (out.argmax(1) == b.y[:b.batch_size]).sum()
      stat[1] += b.batch_size
    net.train()
    dist.all_reduce(stat)
    return float(stat[0] / stat[1].clamp(min=1)), int(stat[1])

  for epoch in range(args.epochs):
    t0 = time.time()
    with Join([model]):                                 # partitions hold different numbers of training seeds
      for b in loader:
        out = model(b.x, b.edge_index, b.num_sampled_nodes, b.num_sampled_edges)[:b.batch_size]
        loss = F.cross_entropy(out, b.y[:b.batch_size])
        opt.zero_grad(); loss.backward(); opt.step()
    print(f'[rank {args.rank}] epoch {epoch} loss {float(loss.detach()):.4f} time {time.time() - t0:.2f}s')
    if test_loader is not None and (epoch + 1) % args.eval_every == 0:
      dist.barrier()
      acc, n = test_accuracy()
      if args.rank == 0:
        print(f'epoch {epoch} test acc {acc:.4f} ({n} nodes)')
  loader.shutdown()
  if test_loader is not None:
    test_loader.shutdown()
  dist.barrier()
  if gd.rpc_is_initialized():
    gd.barrier()
    gd.shutdown_rpc()


if __name__ == '__main__':   # spawned sampling workers re-import this module
  main()
