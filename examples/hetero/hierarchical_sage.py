"""Hierarchical heterogeneous GraphSAGE: layer i only computes the nodes / edges that can still
influence the seeds (per-hop trimming with `num_sampled_nodes` / `num_sampled_edges`), on a MAG-shaped
graph -- counterpart of the reference's examples/hetero/hierarchical_sage.py (PyG trim_to_layer).

  python examples/hetero/hierarchical_sage.py [--no-trim]   # compare the time per epoch with / without trimming
"""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import glt, synthetic_igbh  # noqa: E402
from graphlearn_for_pytorch_b200.models import RGNN  # noqa: E402

if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--papers', type=int, default=8000)
  ap.add_argument('--fanout', default='10,10,5')
  ap.add_argument('--batch', type=int, default=256)
  ap.add_argument('--epochs', type=int, default=2)
  ap.add_argument('--max_steps', type=int, default=-1)
  ap.add_argument('--no-trim', action='store_true')
  a = ap.parse_args()
  cuda = torch.cuda.is_available()
  device = torch.device('cuda', 0) if cuda else torch.device('cpu')
  edges, feats, labels, sizes = synthetic_igbh(a.papers, num_authors=a.papers // 2, feat_dim=64)
  ds = glt.data.Dataset(edge_dir='in')
  ds.init_graph(edges, graph_mode='CUDA' if cuda else 'CPU', num_nodes=sizes)
  ds.init_node_features(feats, with_gpu=cuda, split_ratio=1.0 if cuda else 0.0)
  ds.init_node_labels(labels)
  fan = [int(v) for v in a.fanout.split(',')]
  loader = glt.loader.NeighborLoader(ds, fan, ('paper', torch.randperm(a.papers)[: a.papers // 2]), batch_size=a.batch,
                                     shuffle=True, device=device)
  first = next(iter(loader))
  model = RGNN(list(first.edge_index_dict.keys()), 64, 128, int(labels['paper'].max()) + 1, num_layers=len(fan),
               node_type='paper', model='rsage').to(device)
  opt = torch.optim.Adam(model.parameters(), lr=3e-3)
  for epoch in range(a.epochs):
    t0, correct, seen = time.time(), 0, 0
    for i, b in enumerate(loader):
      if 0 <= a.max_steps <= i:
        break
      bs = b['paper'].batch_size
      if a.no_trim:
        out = model(b.x_dict, b.edge_index_dict)[:bs]
      else:       # hop-wise trimming: the outermost hop is dropped after layer 1, the next after layer 2, ...
        out = model(b.x_dict, b.edge_index_dict, b.num_sampled_nodes, b.num_sampled_edges)[:bs]
      y = b['paper'].y[:bs].to(device)
      loss = F.cross_entropy(out, y)
      opt.zero_grad(); loss.backward(); opt.step()
      correct += int((out.argmax(1) == y).sum()); seen += bs
    if cuda:
      torch.cuda.synchronize()
    print(f'epoch {epoch} ({"full" if a.no_trim else "trimmed"}): loss {float(loss.detach()):.4f} '
          f'train-acc {correct / max(seen, 1):.4f} time {time.time() - t0:.2f}s')
