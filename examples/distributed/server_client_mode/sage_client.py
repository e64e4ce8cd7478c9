"""Training client (server-client mode; counterpart of the reference's
examples/distributed/server_client_mode/sage_supervised_client.py)."""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from common import glt  # noqa: E402,F401
import graphlearn_for_pytorch_b200.distributed as gd  # noqa: E402
from graphlearn_for_pytorch_b200.models import GraphSAGE  # noqa: E402
from graphlearn_for_pytorch_b200.typing import Split  # noqa: E402


def main():
  p = argparse.ArgumentParser()
  p.add_argument('--rank', type=int, default=0)
  p.add_argument('--servers', type=int, default=2)
  p.add_argument('--clients', type=int, default=1)
  p.add_argument('--master-addr', default='127.0.0.1')
  p.add_argument('--master-port', type=int, default=29800)
  p.add_argument('--feat-dim', type=int, default=100)
  p.add_argument('--classes', type=int, default=47)
  args = p.parse_args()

  device = torch.device('cuda', 0) if torch.cuda.is_available() else torch.device('cpu')
  gd.init_client(args.servers, args.clients, args.rank, args.master_addr, args.master_port)
  opts = gd.RemoteDistSamplingWorkerOptions(server_rank=list(range(args.servers)), num_workers=1, worker_concurrency=4,
                                            master_addr=args.master_addr, master_port=args.master_port + 1 + args.rank,
                                            prefetch_size=4)
  loader = gd.DistNeighborLoader(None, [15, 10, 5], Split.train, batch_size=512, shuffle=True, collect_features=True,
                                 to_device=device, worker_options=opts)
  model = GraphSAGE(args.feat_dim, 256, args.classes, 3).to(device)
  opt = torch.optim.Adam(model.parameters(), lr=3e-3)
  for epoch in range(2):
    for b in loader:
      out = model(b.x, b.edge_index, b.num_sampled_nodes, b.num_sampled_edges)[:b.batch_size]
      loss = F.cross_entropy(out, b.y[:b.batch_size])
      opt.zero_grad(); loss.backward(); opt.step()
    print(f'[client {args.rank}] epoch {epoch} loss {float(loss.detach()):.4f}')
  loader.shutdown()
  gd.shutdown_client()


if __name__ == '__main__':   # spawned sampling workers re-import this module
  main()
