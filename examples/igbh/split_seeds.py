"""Fix the train / validation seeds of the labelled paper nodes once (counterpart of the reference's
examples/igbh/split_seeds.py): every training job -- single GPU, multi GPU, distributed -- reads the same files.

  python examples/igbh/split_seeds.py --path /data/igbh --dataset_size tiny --validation_frac 0.05
"""
import argparse
import os.path as osp
import sys

import torch

sys.path.insert(0, osp.dirname(osp.abspath(__file__)))
from dataset import IGBHeteroDataset  # noqa: E402

if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--path', required=True)
  ap.add_argument('--dataset_size', default='tiny')
  ap.add_argument('--random_seed', type=int, default=42)
  ap.add_argument('--num_classes', type=int, default=19, choices=[19, 2983])
  ap.add_argument('--validation_frac', type=float, default=0.005)
  ap.add_argument('--train_frac', type=float, default=0.6)
  a = ap.parse_args()
  ds = IGBHeteroDataset(a.path, a.dataset_size, with_edges=False, use_label_2K=a.num_classes == 2983)
  n = ds.label.numel()
  g = torch.Generator().manual_seed(a.random_seed)
  perm = torch.randperm(n, generator=g)
  n_train, n_val = int(n * a.train_frac), max(int(n * a.validation_frac), 1)
  torch.save(perm[:n_train].clone(), osp.join(ds.base_path, 'train_idx.pt'))
  torch.save(perm[n_train:n_train + n_val].clone(), osp.join(ds.base_path, 'val_idx.pt'))
  print(f'{n_train} training and {n_val} validation seeds written to {ds.base_path}')
