"""Second stage of two-stage partitioning: write the feature files of ONE partition from the full feature
matrices, using the partition books stored by `partition.py --with_feature 0` (counterpart of the reference's
examples/igbh/build_partition_feature.py).  Run once per partition, typically on the machine that will train it.

  python examples/igbh/build_partition_feature.py --src_path /data/igbh --dst_path /data/igbh_parts --partition_idx 0
"""
import argparse
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
  ap.add_argument('--in_memory', type=int, default=0)
  ap.add_argument('--partition_idx', type=int, default=0)
  ap.add_argument('--chunk_size', type=int, default=10000)
  ap.add_argument('--use_fp16', action='store_true')
  a = ap.parse_args()
  ds = IGBHeteroDataset(a.src_path, a.dataset_size, in_memory=bool(a.in_memory), with_edges=False, use_fp16=a.use_fp16)
  glt.partition.build_partition_feature(a.dst_path, a.partition_idx, chunk_size=a.chunk_size, node_feat=ds.feat_dict,
                                        node_feat_dtype=torch.float16 if a.use_fp16 else torch.float32)
  print(f'features of partition {a.partition_idx} written under {a.dst_path}')
