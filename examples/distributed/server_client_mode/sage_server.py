"""Sampling server (server-client mode; counterpart of the reference's
examples/distributed/server_client_mode/sage_supervised_server.py)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from common import glt  # noqa: E402,F401
import graphlearn_for_pytorch_b200.distributed as gd  # noqa: E402


def main():
  p = argparse.ArgumentParser()
  p.add_argument('--root', required=True)
  p.add_argument('--rank', type=int, required=True)
  p.add_argument('--servers', type=int, default=2)
  p.add_argument('--clients', type=int, default=1)
  p.add_argument('--master-addr', default='127.0.0.1')
  p.add_argument('--master-port', type=int, default=29800)
  args = p.parse_args()

  ds = gd.DistDataset()
  ds.load(args.root, args.rank, graph_mode='CPU', feature_with_gpu=False,
          whole_node_label_file=os.path.join(args.root, 'labels.pt'))
  train = torch.load(os.path.join(args.root, 'train_idx.pt'))
  own = train[ds.node_pb[train] == args.rank]
  ds.init_node_split((own, own[:0], own[:0]))
  gd.init_server(args.servers, args.rank, ds, args.master_addr, args.master_port, num_clients=args.clients)
  gd.wait_and_shutdown_server()


if __name__ == '__main__':   # spawned sampling workers re-import this module
  main()
