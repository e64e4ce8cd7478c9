"""HGT on an OGB-MAG-shaped heterogeneous graph through the hetero `NeighborLoader`.

Counterpart of the reference's examples/hetero/train_hgt_mag.py (PyG HGTConv, hidden 64, 2 heads, 2 layers,
fanout [10, 10], batch 1024, ZERO_COPY topology + 20 % of the features on the GPU): the model is this repo's
dependency-free `models.HGT`, the graph is synthetic with the MAG schema (no network access; swap
`synthetic_mag` for the real tensors).  On a GPU the loader samples through the native grouped hetero arena
(one launch per hop over all seven relations).
"""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import glt, synthetic_mag  # noqa: E402
from graphlearn_for_pytorch_b200.models import HGT  # noqa: E402


def build_dataset(args, cuda, device_index=0):
  edges, feats, labels, sizes = synthetic_mag(args.papers)
  ds = glt.data.Dataset()
  ds.init_graph(edges, graph_mode=('ZERO_COPY' if args.zero_copy else 'CUDA') if cuda else 'CPU', num_nodes=sizes,
                device=device_index)
  ds.init_node_features(feats, split_ratio=args.split_ratio if cuda else 0.0, with_gpu=cuda, device=device_index)
  ds.init_node_labels(labels)
  n = sizes['paper']
  perm = torch.randperm(n, generator=torch.Generator().manual_seed(1))
  return ds, sizes, labels, perm[: n // 2], perm[n // 2: n // 2 + n // 10]


def make_model(ds, sizes, labels, loader, args, device):
  first = next(iter(loader))
  return HGT(list(sizes.keys()), list(first.edge_index_dict.keys()), 128, args.hidden, int(labels['paper'].max()) + 1,
             num_layers=2, heads=args.heads, node_type='paper').to(device)


def train_epoch(model, loader, opt, device):
  model.train()
  tot = seen = 0
  for b in loader:
    bs = b['paper'].batch_size
    out = model(b.x_dict, b.edge_index_dict)[:bs]
    loss = F.cross_entropy(out, b['paper'].y[:bs].to(device))
    opt.zero_grad(); loss.backward(); opt.step()
    tot += float(loss.detach()) * bs; seen += bs
  return tot / max(seen, 1)


@torch.no_grad()
def evaluate(model, loader, device):
  model.eval()
  correct = seen = 0
  for b in loader:
    bs = b['paper'].batch_size
    pred = model(b.x_dict, b.edge_index_dict)[:bs].argmax(-1)
    correct += int((pred == b['paper'].y[:bs].to(device)).sum()); seen += bs
  return correct / max(seen, 1)


def parse():
  p = argparse.ArgumentParser()
  p.add_argument('--papers', type=int, default=30_000)
  p.add_argument('--epochs', type=int, default=3)
  p.add_argument('--batch', type=int, default=1024)
  p.add_argument('--hidden', type=int, default=64)
  p.add_argument('--heads', type=int, default=2)
  p.add_argument('--split-ratio', type=float, default=0.2)
  p.add_argument('--zero-copy', action='store_true', help='keep the topology in pinned host memory (reference default)')
  return p.parse_args()


if __name__ == '__main__':
  args = parse()
  cuda = torch.cuda.is_available()
  device = torch.device('cuda', 0) if cuda else torch.device('cpu')
  ds, sizes, labels, train_idx, val_idx = build_dataset(args, cuda)
  train_loader = glt.loader.NeighborLoader(ds, [10, 10], ('paper', train_idx), batch_size=args.batch, shuffle=True,
                                           device=device)
  val_loader = glt.loader.NeighborLoader(ds, [10, 10], ('paper', val_idx), batch_size=args.batch, device=device)
  model = make_model(ds, sizes, labels, train_loader, args, device)
  opt = torch.optim.Adam(model.parameters(), lr=0.01)
  for epoch in range(1, args.epochs + 1):
    t0 = time.time()
    loss = train_epoch(model, train_loader, opt, device)
    acc = evaluate(model, val_loader, device)
    print(f'Epoch: {epoch:02d}, Loss: {loss:.4f}, Val: {acc:.4f}, Time: {time.time() - t0:.2f}s')
