"""Pre-compress the IGBH edge lists into CSC / CSR once, so that training jobs skip the COO -> CSR
conversion at start-up (counterpart of the reference's examples/igbh/compress_graph.py).

  python examples/igbh/compress_graph.py --path /data/igbh --dataset_size tiny --layout CSC [--use_fp16]
Writes <processed>/<layout>/<etype>/{compressed_0.pt, compressed_1.pt, edge_ids.pt}:
  CSC: compressed_0 = row indices, compressed_1 = column pointer   (what edge_dir='in' sampling reads)
  CSR: compressed_0 = row pointer, compressed_1 = column indices   (edge_dir='out')
"""
import argparse
import os
import os.path as osp
import sys
import time

import torch

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
sys.path.insert(0, osp.dirname(osp.abspath(__file__)))
from common import glt  # noqa: E402
from dataset import IGBHeteroDataset, etype_dir, float2half  # noqa: E402

if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--path', required=True)
  ap.add_argument('--dataset_size', default='tiny')
  ap.add_argument('--layout', default='CSC', choices=['CSC', 'CSR'])
  ap.add_argument('--use_fp16', action='store_true', help='also convert the feature files to fp16')
  a = ap.parse_args()
  ds = IGBHeteroDataset(a.path, a.dataset_size, layout='COO')
  n = ds.num_nodes
  t0 = time.time()
  for et, ei in ds.edge_dict.items():
    topo = glt.data.Topology(ei, layout=a.layout, num_nodes=n[et[2]] if a.layout == "CSC" else n[et[0]])  # native op
    d = osp.join(ds.base_path, a.layout, etype_dir(et))
    os.makedirs(d, exist_ok=True)
    first, second = (topo.indices, topo.indptr) if a.layout == 'CSC' else (topo.indptr, topo.indices)
    torch.save(first, osp.join(d, 'compressed_0.pt'))
    torch.save(second, osp.join(d, 'compressed_1.pt'))
    torch.save(topo.edge_ids, osp.join(d, 'edge_ids.pt'))
    print(f'{et}: {ei.shape[1]} edges -> {a.layout}')
  if a.use_fp16:
    float2half(ds.base_path)
  print(f'compressed {len(ds.edge_dict)} relations in {time.time() - t0:.2f} s')
