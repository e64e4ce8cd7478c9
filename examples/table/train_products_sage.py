"""GraphSAGE trained from *tables* through `TableDataset` (counterpart of the reference's
examples/pai/ogbn_products/train_products_sage.py, which reads ODPS tables via common_io; here any
parquet / CSV / pyarrow table works).

  python examples/table/data_preprocess.py --out /tmp/products_tables
  python examples/table/train_products_sage.py --tables /tmp/products_tables --epochs 2
"""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import glt  # noqa: E402
from graphlearn_for_pytorch_b200.models import GraphSAGE  # noqa: E402

if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--tables', required=True)
  ap.add_argument('--epochs', type=int, default=2)
  ap.add_argument('--batch_size', type=int, default=512)
  ap.add_argument('--fanout', default='10,5')
  ap.add_argument('--max_steps', type=int, default=-1)
  a = ap.parse_args()
  cuda = torch.cuda.is_available()
  device = torch.device('cuda', 0) if cuda else torch.device('cpu')
  ds = glt.data.TableDataset()
  ds.load(edge_tables={('item', 'link', 'item'): os.path.join(a.tables, 'edge.parquet')},
          node_tables={'item': os.path.join(a.tables, 'node.parquet')},
          graph_mode='CUDA' if cuda else 'CPU', sort_func=glt.data.sort_by_in_degree,
          split_ratio=0.2 if cuda else 0.0, directed=False, label='label', weight_col='weight')
  n = ds.node_labels.numel()
  train_idx = torch.randperm(n)[: n // 2]
  fan = [int(v) for v in a.fanout.split(',')]
  loader = glt.loader.NeighborLoader(ds, fan, train_idx, batch_size=a.batch_size, shuffle=True, device=device)
  model = GraphSAGE(ds.node_features.shape[1], 128, int(ds.node_labels.max()) + 1, len(fan)).to(device)
  opt = torch.optim.Adam(model.parameters(), lr=3e-3)
  for epoch in range(a.epochs):
    t0, correct, seen = time.time(), 0, 0
    for i, b in enumerate(loader):
      if 0 <= a.max_steps <= i:
        break
      out = model(b.x, b.edge_index, b.num_sampled_nodes, b.num_sampled_edges)[:b.batch_size]
      y = b.y[:b.batch_size]
      loss = F.cross_entropy(out, y)
      opt.zero_grad(); loss.backward(); opt.step()
      correct += int((out.argmax(1) == y).sum()); seen += y.numel()
    print(f'epoch {epoch}: loss {float(loss.detach()):.4f} train-acc {correct / max(seen, 1):.4f} ({time.time() - t0:.1f}s)')
