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


def load_mag_from_disk(root, name='ogbn-mag'):
  """ogbn-mag from its OGB directory: reverse relations are added (what T.ToUndirected(merge=True) does in the
  reference script) and node types that ship without features get the mean of their featured neighbours' rows,
  propagated until every type has some (author <- papers written, institution <- authors, field <- papers); the
  reference substitutes pre-trained metapath2vec vectors there (examples/hetero/train_hgt_mag.py:33-36)."""
  d = glt.utils.load_ogb_hetero_dataset(root, name)
  edges = dict(d['edge_index'])
  for (s, r, t), ei in list(edges.items()):
    if s == t:
      edges[(s, r, t)] = torch.cat([ei, ei.flip(0)], 1)
    elif (t, f'rev_{r}', s) not in edges:
      edges[(t, f'rev_{r}', s)] = ei.flip(0)
  sizes, feats = dict(d['num_nodes']), dict(d['x'])
  while len(feats) < len(sizes):
    progressed = False
    for (s, r, t), ei in edges.items():
      if s in feats and t not in feats:
        acc = torch.zeros(sizes[t], feats[s].shape[1]).index_add_(0, ei[1], feats[s][ei[0]])
        deg = torch.zeros(sizes[t]).index_add_(0, ei[1], torch.ones(ei.shape[1]))
        feats[t] = acc / deg.clamp(min=1).unsqueeze(1)
        progressed = True
    if not progressed:
      raise ValueError(f'node types without features and without a featured neighbour type: '
                       f'{sorted(set(sizes) - set(feats))}')
  labels = {t: v.to(torch.int64) for t, v in d['y'].items()}
  return edges, feats, labels, sizes, d['split']


def build_dataset(args, cuda, device_index=0):
  split = None
  if getattr(args, 'root', None):
    edges, feats, labels, sizes, split = load_mag_from_disk(args.root, args.dataset)
  else:
    edges, feats, labels, sizes = synthetic_mag(args.papers)
  ds = glt.data.Dataset()
  ds.init_graph(edges, graph_mode=('ZERO_COPY' if args.zero_copy else 'CUDA') if cuda else 'CPU', num_nodes=sizes,
                device=device_index)
  ds.init_node_features(feats, split_ratio=args.split_ratio if cuda else 0.0, with_gpu=cuda, device=device_index)
  ds.init_node_labels(labels)
  if split and 'train' in split and 'paper' in split['train']:
    return ds, sizes, labels, split['train']['paper'], split.get('valid', split['train'])['paper']
  n = sizes['paper']
  perm = torch.randperm(n, generator=torch.Generator().manual_seed(1))
  return ds, sizes, labels, perm[: n // 2], perm[n // 2: n // 2 + n // 10]


def make_model(ds, sizes, labels, loader, args, device):
  first = next(iter(loader))
  in_dim = ds.node_features['paper'].shape[1]
  return HGT(list(sizes.keys()), list(first.edge_index_dict.keys()), in_dim, args.hidden, int(labels['paper'].max()) + 1,
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
  p.add_argument('--root', default=None, help='directory that contains ogbn_mag/ (OGB raw layout); default: synthetic')
  p.add_argument('--dataset', default='ogbn-mag')
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
