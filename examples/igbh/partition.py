"""Partition the IGBH graph for distributed training (counterpart of the reference's examples/igbh/partition.py).

  python examples/igbh/partition.py --src_path /data/igbh --dst_path /data/igbh_parts --num_partitions 2 \
      [--with_feature 0] [--edge_assign_strategy by_dst] [--layout CSC] [--graph_caching 1] [--data_precision fp16]
With --with_feature 0 only topology + partition books are written; every machine then builds its own feature
files with build_partition_feature.py (two-stage partitioning: the full feature matrix never has to fit on the
partitioning host together with the graph).
"""
import argparse
import os
import os.path as osp
import sys

import torch

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
sys.path.insert(0, osp.dirname(osp.abspath(__file__)))
from common import glt  # noqa: E402
from dataset import IGBHeteroDataset  # noqa: E402

if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--src_path', required=True)
  ap.add_argument('--dst_path', required=True)
  ap.add_argument('--dataset_size', default='tiny')
  ap.add_argument('--num_classes', type=int, default=19, choices=[19, 2983])
  ap.add_argument('--in_memory', type=int, default=1)
  ap.add_argument('--num_partitions', type=int, default=2)
  ap.add_argument('--chunk_size', type=int, default=10000)
  ap.add_argument('--edge_assign_strategy', default='by_src', choices=['by_src', 'by_dst'])
  ap.add_argument('--with_feature', type=int, default=1)
  ap.add_argument('--graph_caching', type=int, default=0)
  ap.add_argument('--data_precision', default='fp32', choices=['fp32', 'fp16'])
  ap.add_argument('--layout', default='COO', choices=['COO', 'CSC', 'CSR'],
                  help='additionally store every partition pre-compressed in this layout')
  a = ap.parse_args()
  ds = IGBHeteroDataset(a.src_path, a.dataset_size, in_memory=bool(a.in_memory), use_label_2K=a.num_classes == 2983,
                        use_fp16=a.data_precision == 'fp16')
  dtype = torch.float16 if a.data_precision == 'fp16' else torch.float32
  part = glt.partition.RandomPartitioner(
    a.dst_path, a.num_partitions, ds.num_nodes, ds.edge_dict,
    node_feat=ds.feat_dict if a.with_feature else None, node_feat_dtype=dtype,
    edge_assign_strategy=a.edge_assign_strategy, chunk_size=a.chunk_size)
  part.partition(with_feature=bool(a.with_feature), graph_caching=bool(a.graph_caching))
  torch.save(ds.label, osp.join(a.dst_path, 'label.pt'))
  for name in ('train_idx', 'val_idx'):
    idx = getattr(ds, name)
    if idx is not None:
      torch.save(idx, osp.join(a.dst_path, f'{name}.pt'))
  if a.layout != 'COO':      # per-partition compressed copies (skips COO -> CSR/CSC at load time)
    from graphlearn_for_pytorch_b200.partition.base import load_partition
    for p in range(a.num_partitions):
      _, _, graph, _, _, _, _ = load_partition(a.dst_path, p)
      for et, g in graph.items():
        n = ds.num_nodes[et[2] if a.layout == 'CSC' else et[0]]
        topo = glt.data.Topology(torch.stack(list(g.edge_index)), g.eids, layout=a.layout, num_nodes=n)
        d = osp.join(a.dst_path, f'part{p}', a.layout, '__'.join(et))
        os.makedirs(d, exist_ok=True)
        first, second = (topo.indices, topo.indptr) if a.layout == 'CSC' else (topo.indptr, topo.indices)
        torch.save((first, second, topo.edge_ids), osp.join(d, 'compressed.pt'))
  print(f'{a.num_partitions} partitions written to {a.dst_path} (features: {bool(a.with_feature)}, {a.data_precision})')
