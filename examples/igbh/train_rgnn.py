"""Single-process R-GNN training on the on-disk IGBH dataset (GPU if present; counterpart of the reference's
examples/igbh/train_rgnn_multi_gpu.py for one device -- the multi-GPU/device-resident path of this framework is
`examples/multi_gpu/train_sage_p2p.py` / `bench.py`, the distributed one is dist_train_rgnn.py).

  python examples/igbh/train_rgnn.py --path /data/igbh --dataset_size tiny [--layout CSC] [--use_fp16]
"""
import argparse
import os.path as osp
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
sys.path.insert(0, osp.dirname(osp.abspath(__file__)))
from common import glt  # noqa: E402
from dataset import IGBHeteroDataset  # noqa: E402
from graphlearn_for_pytorch_b200.models import RGNN  # noqa: E402
from mlperf_logging_utils import get_mlperf_logger  # noqa: E402

if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--path', required=True)
  ap.add_argument('--dataset_size', default='tiny')
  ap.add_argument('--layout', default='COO', choices=['COO', 'CSC', 'CSR'])
  ap.add_argument('--use_fp16', action='store_true')
  ap.add_argument('--model', default='rgat', choices=['rgat', 'rsage', 'rgcn'])
  ap.add_argument('--fan_out', default='15,10,5')
  ap.add_argument('--batch_size', type=int, default=512)
  ap.add_argument('--epochs', type=int, default=2)
  ap.add_argument('--max_steps', type=int, default=-1)
  a = ap.parse_args()
  cuda = torch.cuda.is_available()
  device = torch.device('cuda', 0) if cuda else torch.device('cpu')
  log = get_mlperf_logger()
  log.start('INIT')
  igbh = IGBHeteroDataset(a.path, a.dataset_size, layout=a.layout, use_fp16=a.use_fp16)
  edge_dir = 'out' if a.layout == 'CSR' else 'in'
  ds = glt.data.Dataset(edge_dir=edge_dir)
  ds.init_graph(igbh.edge_dict, layout=a.layout, graph_mode='CUDA' if cuda else 'CPU', num_nodes=igbh.num_nodes)
  ds.init_node_features(igbh.feat_dict, with_gpu=cuda, split_ratio=1.0 if cuda else 0.0,
                        dtype=torch.float16 if a.use_fp16 else torch.float32)
  ds.init_node_labels({'paper': igbh.label})
  assert igbh.train_idx is not None, 'run split_seeds.py first'
  fan = [int(v) for v in a.fan_out.split(',')]
  loader = glt.loader.NeighborLoader(ds, fan, ('paper', igbh.train_idx), batch_size=a.batch_size, shuffle=True,
                                     device=device)
  val_loader = glt.loader.NeighborLoader(ds, fan, ('paper', igbh.val_idx), batch_size=a.batch_size, device=device)
  first = next(iter(loader))
  in_dim = next(iter(first.x_dict.values())).shape[1]
  model = RGNN(list(first.edge_index_dict.keys()), in_dim, 128, igbh.num_classes, num_layers=len(fan), node_type='paper',
               model=a.model).to(device)
  opt = torch.optim.Adam(model.parameters(), lr=1e-3)
  log.end('INIT')
  log.start('RUN')
  for epoch in range(a.epochs):
    t0 = time.time()
    for i, b in enumerate(loader):
      if 0 <= a.max_steps <= i:
        break
      bs = b['paper'].batch_size
      x = {k: v.float() for k, v in b.x_dict.items()}
      out = model(x, b.edge_index_dict, b.num_sampled_nodes, b.num_sampled_edges)[:bs]
      loss = F.cross_entropy(out, b['paper'].y[:bs].to(device))
      opt.zero_grad(); loss.backward(); opt.step()
    model.eval()
    correct = total = 0
    with torch.no_grad():
      for b in val_loader:
        bs = b['paper'].batch_size
        x = {k: v.float() for k, v in b.x_dict.items()}
        out = model(x, b.edge_index_dict, b.num_sampled_nodes, b.num_sampled_edges)[:bs]
        correct += int((out.argmax(1) == b['paper'].y[:bs].to(device)).sum()); total += bs
    model.train()
    log.event('EVAL_ACCURACY', correct / max(total, 1), {'epoch_num': epoch})
    print(f'epoch {epoch}: loss {float(loss.detach()):.4f} val-acc {correct / max(total, 1):.4f} ({time.time() - t0:.1f}s)')
  log.end('RUN')
