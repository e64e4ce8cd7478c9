"""Distributed GraphSAGE from table *slices* (counterpart of the reference's
examples/pai/ogbn_products/dist_train_products_sage.py): every worker reads its slice of the node / edge
tables, the workers partition the graph collectively online (`DistTableDataset` -> DistRandomPartitioner over
RPC) and then train with `DistNeighborLoader` + DDP.

  python examples/table/data_preprocess.py --out /tmp/products_tables
  for r in 0 1; do python examples/table/dist_train_products_sage.py --tables /tmp/products_tables --rank $r --world 2 & done
"""
import argparse
import os
import sys

import pyarrow.parquet as pq
import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import glt  # noqa: E402,F401
import graphlearn_for_pytorch_b200.distributed as gd  # noqa: E402
from graphlearn_for_pytorch_b200.models import GraphSAGE  # noqa: E402


def table_slice(path, rank, world):
  """What a table reader with (slice_id, slice_count) returns: a contiguous share of the rows."""
  t = pq.read_table(path)
  per = (t.num_rows + world - 1) // world
  return t.slice(rank * per, per)


if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--tables', required=True)
  ap.add_argument('--rank', type=int, required=True)
  ap.add_argument('--world', type=int, default=2)
  ap.add_argument('--master_addr', default='127.0.0.1')
  ap.add_argument('--master_port', type=int, default=29860)
  ap.add_argument('--epochs', type=int, default=1)
  ap.add_argument('--max_steps', type=int, default=-1)
  ap.add_argument('--out', default=None, help='shared directory for the online partitions (default: temp dir of rank 0)')
  a = ap.parse_args()
  cuda = torch.cuda.is_available()
  device = torch.device('cuda', a.rank % max(torch.cuda.device_count(), 1)) if cuda else torch.device('cpu')
  os.environ.setdefault('MASTER_ADDR', a.master_addr)
  os.environ.setdefault('MASTER_PORT', str(a.master_port))
  dist.init_process_group('nccl' if cuda else 'gloo', rank=a.rank, world_size=a.world)
  gd.init_worker_group(a.world, a.rank)
  gd.init_rpc(a.master_addr, a.master_port + 1)
  num_nodes = pq.read_metadata(os.path.join(a.tables, 'node.parquet')).num_rows
  ds = gd.DistTableDataset()
  kw = {'output_dir': a.out} if a.out else {}
  ds.load(num_nodes, {None: table_slice(os.path.join(a.tables, 'edge.parquet'), a.rank, a.world)},
          {None: table_slice(os.path.join(a.tables, 'node.parquet'), a.rank, a.world)},
          graph_mode='CUDA' if cuda else 'CPU', feature_with_gpu=cuda, **kw)
  g = torch.Generator().manual_seed(0)
  train = torch.randperm(num_nodes, generator=g)[: num_nodes // 2]
  per = train.numel() // a.world
  train = train[a.rank * per:(a.rank + 1) * per]          # equal shares: same number of DDP steps per rank
  loader = gd.DistNeighborLoader(ds, [10, 5], train, batch_size=256, shuffle=True, collect_features=True,
                                 to_device=device, worker_options=gd.CollocatedDistSamplingWorkerOptions(
                                   master_addr=a.master_addr, master_port=a.master_port + 1))
  n_cls = int(ds.node_labels.max()) + 1
  model = torch.nn.parallel.DistributedDataParallel(GraphSAGE(ds.node_features.shape[1], 128, n_cls, 2).to(device))
  opt = torch.optim.Adam(model.parameters(), lr=3e-3)
  for epoch in range(a.epochs):
    for i, b in enumerate(loader):
      if 0 <= a.max_steps <= i:
        continue
      out = model(b.x, b.edge_index, b.num_sampled_nodes, b.num_sampled_edges)[:b.batch_size]
      loss = F.cross_entropy(out, b.y[:b.batch_size])
      opt.zero_grad(); loss.backward(); opt.step()
    print(f'[rank {a.rank}] epoch {epoch} loss {float(loss.detach()):.4f}', flush=True)
  loader.shutdown()
  dist.barrier()
  if gd.rpc_is_initialized():
    gd.barrier()
    gd.shutdown_rpc()
  dist.destroy_process_group()
