"""Distributed unsupervised GraphSAGE with ZeRO-1 optimizer-state sharding and uneven-input Join --
counterpart of the reference's examples/distributed/dist_sage_unsup/dist_sage_unsup.py:28,143-149.
Launch like dist_train_sage.py (one process per partition)."""
import argparse
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch.distributed.algorithms.join import Join
from torch.distributed.optim import ZeroRedundancyOptimizer

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import glt  # noqa: E402,F401
import graphlearn_for_pytorch_b200.distributed as gd  # noqa: E402
from graphlearn_for_pytorch_b200.models import GraphSAGE  # noqa: E402
from graphlearn_for_pytorch_b200.sampler import NegativeSampling  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument('--root', required=True)
p.add_argument('--rank', type=int, required=True)
p.add_argument('--world', type=int, default=2)
p.add_argument('--master-addr', default='127.0.0.1')
p.add_argument('--master-port', type=int, default=29750)
args = p.parse_args()
cuda = torch.cuda.is_available()
device = torch.device('cuda', args.rank % max(torch.cuda.device_count(), 1)) if cuda else torch.device('cpu')
os.environ.setdefault('MASTER_ADDR', args.master_addr)
os.environ.setdefault('MASTER_PORT', str(args.master_port))
dist.init_process_group('nccl' if cuda else 'gloo', rank=args.rank, world_size=args.world)
gd.init_worker_group(args.world, args.rank)
ds = gd.DistDataset().load(args.root, args.rank, graph_mode='CUDA' if cuda else 'CPU', feature_with_gpu=cuda,
                           device=device.index)
rows, cols, _, _ = ds.graph.topo.to_coo()
opts = gd.CollocatedDistSamplingWorkerOptions(master_addr=args.master_addr, master_port=args.master_port + 1)
loader = gd.DistLinkNeighborLoader(ds, [10, 5], batch_size=512, edge_label_index=torch.stack([rows, cols])[:, :20000],
                                   neg_sampling=NegativeSampling('binary', 1), shuffle=True, collect_features=True,
                                   to_device=device, worker_options=opts)
model = torch.nn.parallel.DistributedDataParallel(GraphSAGE(ds.node_features.shape[1], 128, 64, 2).to(device))
opt = ZeroRedundancyOptimizer(model.parameters(), optimizer_class=torch.optim.Adam, lr=3e-3)
for epoch in range(2):
  with Join([model, opt]):        # partitions own different numbers of links
    for b in loader:
      h = model(b.x, b.edge_index)
      logit = (h[b.edge_label_index[1]] * h[b.edge_label_index[0]]).sum(-1)
      loss = F.binary_cross_entropy_with_logits(logit, b.edge_label.float())
      opt.zero_grad(); loss.backward(); opt.step()
  print(f'[rank {args.rank}] epoch {epoch} loss {float(loss.detach()):.4f}')
loader.shutdown()
