"""Multi-process (one process per GPU, DDP) HGT training on a MAG-shaped heterogeneous graph.

Counterpart of the reference's examples/hetero/train_hgt_mag_mp.py: the dataset is built once, shared with the
spawned trainers through IPC handles (`Dataset.share_ipc()`: GPU shards travel as CUDA-IPC handles opened on the
consumer's device, host parts as shared memory), every rank trains on its slice of the seeds with a hetero
`NeighborLoader`, gradients are averaged by DistributedDataParallel.  Falls back to gloo + CPU when no GPU is
visible so the script stays runnable anywhere.
"""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import glt  # noqa: E402
from train_hgt_mag import build_dataset, evaluate, make_model, parse, train_epoch  # noqa: E402


def run(rank, world, ds, sizes, labels, train_idx, val_idx, args, port):
  cuda = torch.cuda.is_available()
  os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
  dist.init_process_group('nccl' if cuda else 'gloo', rank=rank, world_size=world)
  if cuda:
    torch.cuda.set_device(rank)
  device = torch.device('cuda', rank) if cuda else torch.device('cpu')
  torch.manual_seed(42)
  mine = train_idx.split((train_idx.numel() + world - 1) // world)[rank]
  train_loader = glt.loader.NeighborLoader(ds, [10, 10], ('paper', mine), batch_size=args.batch, shuffle=True,
                                           drop_last=True, device=device)
  val_loader = glt.loader.NeighborLoader(ds, [10, 10], ('paper', val_idx), batch_size=args.batch, device=device)
  model = make_model(ds, sizes, labels, train_loader, args, device)
  model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[rank] if cuda else None,
                                                    find_unused_parameters=True)
  opt = torch.optim.Adam(model.parameters(), lr=0.01)
  for epoch in range(1, args.epochs + 1):
    t0 = time.time()
    loss = train_epoch(model, train_loader, opt, device)
    dist.barrier()
    if rank == 0:
      acc = evaluate(model.module, val_loader, device)
      print(f'Epoch: {epoch:02d}, Loss: {loss:.4f}, Val: {acc:.4f}, Time: {time.time() - t0:.2f}s', flush=True)
    dist.barrier()
  dist.destroy_process_group()


if __name__ == '__main__':
  args = parse()
  cuda = torch.cuda.is_available()
  world = max(torch.cuda.device_count(), 1) if cuda else 2
  ds, sizes, labels, train_idx, val_idx = build_dataset(args, cuda)
  ds.share_ipc()
  train_idx.share_memory_(); val_idx.share_memory_()
  import socket
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
  mp.spawn(run, args=(world, ds, sizes, labels, train_idx, val_idx, args, port), nprocs=world, join=True)
