"""R-GNN (R-SAGE / R-GCN / R-GAT) on an IGBH-shaped heterogeneous graph.

Counterpart of the reference's MLPerf-GNN example examples/igbh/train_rgnn_multi_gpu.py
(hetero NeighborLoader, fanout 15,10,5, RGNN with trim_to_layer, checkpoints, mllog events).
"""
import argparse
import time

import torch
import torch.nn.functional as F

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import glt, synthetic_igbh  # noqa: E402
from graphlearn_for_pytorch_b200.models import HGT, RGNN  # noqa: E402
from graphlearn_for_pytorch_b200.utils import EventLogger, load_ckpt, save_ckpt  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument('--model', default='rsage', choices=['rsage', 'rgcn', 'rgat', 'hgt'])
p.add_argument('--papers', type=int, default=20_000)
p.add_argument('--fanout', default='15,10')
p.add_argument('--batch', type=int, default=512)
p.add_argument('--epochs', type=int, default=2)
p.add_argument('--edge-dir', default='in', choices=['in', 'out'])
p.add_argument('--ckpt-dir', default=None)
args = p.parse_args()

cuda = torch.cuda.is_available()
device = torch.device('cuda', 0) if cuda else torch.device('cpu')
log = EventLogger()
log.start('INIT')
edges, feats, labels, sizes = synthetic_igbh(args.papers)
ds = glt.data.Dataset(edge_dir=args.edge_dir)
ds.init_graph(edges, graph_mode='CUDA' if cuda else 'CPU', num_nodes=sizes)
ds.init_node_features(feats, with_gpu=cuda, split_ratio=1.0 if cuda else 0.0, dtype=torch.float32)
ds.init_node_labels(labels)
fan = [int(v) for v in args.fanout.split(',')]
train_idx = torch.randperm(args.papers)[: args.papers // 2]
loader = glt.loader.NeighborLoader(ds, fan, ('paper', train_idx), batch_size=args.batch, shuffle=True,
                                   drop_last=True, device=device)
first = next(iter(loader))
if args.model == 'hgt':
  model = HGT(list(sizes.keys()), list(first.edge_index_dict.keys()), 128, 256, int(labels['paper'].max()) + 1,
              num_layers=len(fan), heads=4, node_type='paper').to(device)
else:
  model = RGNN(list(first.edge_index_dict.keys()), 128, 256, int(labels['paper'].max()) + 1, num_layers=len(fan),
               node_type='paper', model=args.model).to(device)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
start_epoch = 0
if args.ckpt_dir:
  start_epoch = max(0, load_ckpt(0, args.ckpt_dir, model, opt) + 1)
log.end('INIT')
log.start('RUN')
for epoch in range(start_epoch, args.epochs):
  t0, correct, seen = time.time(), 0, 0
  for b in loader:
    out = (model(b.x_dict, b.edge_index_dict) if args.model == 'hgt' else
           model(b.x_dict, b.edge_index_dict, b.num_sampled_nodes, b.num_sampled_edges))[:b['paper'].batch_size]
    tgt = b['paper'].y[:b['paper'].batch_size].to(device)
    loss = F.cross_entropy(out, tgt)
    opt.zero_grad(); loss.backward(); opt.step()
    correct += int((out.argmax(1) == tgt).sum()); seen += tgt.numel()
  log.event('EVAL_ACCURACY', correct / seen, {'epoch': epoch})
  print(f'epoch {epoch}: loss {float(loss.detach()):.4f} train-acc {correct / seen:.4f} time {time.time() - t0:.2f}s')
  if args.ckpt_dir:
    save_ckpt(0, args.ckpt_dir, model, opt, epoch, extra=loader.state_dict())
log.end('RUN')
