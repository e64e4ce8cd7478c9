"""Share one Feature (hot rows in HBM, cold rows in shared pinned host memory) with spawned
processes through its ForkingPickler reducer -- counterpart of the reference's examples/feature_mp.py."""
import torch
import torch.multiprocessing as mp

from common import glt


def worker(rank, feature, n):
  torch.cuda.set_device(rank % max(torch.cuda.device_count(), 1)) if torch.cuda.is_available() else None
  ids = torch.randint(0, n, (8,))
  rows = feature[ids.cuda()] if torch.cuda.is_available() else feature.cpu_get(ids)
  print(f'[proc {rank}] ids {ids.tolist()} -> first column {rows[:, 0].tolist()}')


if __name__ == '__main__':
  n = 10_000
  feat = torch.arange(n, dtype=torch.float32).unsqueeze(1).repeat(1, 16)
  devices = list(range(torch.cuda.device_count())) or [0]
  feature = glt.data.Feature(feat, split_ratio=0.3, device_group_list=[glt.data.DeviceGroup(0, devices)], device=0,
                             with_gpu=torch.cuda.is_available())
  mp.spawn(worker, args=(feature, n), nprocs=2, join=True)
