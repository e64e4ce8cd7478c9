"""Shared-memory sample channel throughput: one producer process -> one consumer, SampleMessage-shaped dicts
(ids, edge index, a feature block), this library vs the unmodified reference (baseline/_ref) on the same container.

  python benchmarks/bench_shm_channel.py ours | reference [--mb 16] [--msgs 200]
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(impl):
  if impl == 'reference':
    sys.path.insert(0, os.path.join(ROOT, 'baseline', 'shims'))
    sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))
    import graphlearn_torch as glt
  else:
    sys.path.insert(0, ROOT)
    import graphlearn_for_pytorch_b200 as glt
  return glt


def message(mb):
  rows = max(int(mb * 2 ** 20 / (128 * 4 + 8 + 32)), 16)
  return {'ids': torch.arange(rows), 'rows': torch.arange(2 * rows), 'cols': torch.arange(2 * rows),
          'nfeats': torch.randn(rows, 128)}


def producer(impl, ch, mb, msgs):
  load(impl)
  m = message(mb)
  for _ in range(msgs):
    ch.send(m)


if __name__ == '__main__':
  p = argparse.ArgumentParser()
  p.add_argument('impl', choices=['ours', 'reference'])
  p.add_argument('--mb', type=float, default=16.0)
  p.add_argument('--msgs', type=int, default=200)
  a = p.parse_args()
  glt = load(a.impl)
  mp.set_start_method('spawn', force=True)
  ch = glt.channel.ShmChannel(capacity=8, shm_size='512MB')
  nbytes = sum(t.numel() * t.element_size() for t in message(a.mb).values())
  pr = mp.Process(target=producer, args=(a.impl, ch, a.mb, a.msgs + 3))
  pr.start()
  for _ in range(3):
    ch.recv()
  t = time.time()
  chk = 0
  for _ in range(a.msgs):
    chk += int(ch.recv()['ids'][-1])
  dt = time.time() - t
  pr.join()
  print(json.dumps({'impl': a.impl, 'msg_MB': nbytes / 2 ** 20, 'msgs_per_s': a.msgs / dt,
                    'GB_per_s': nbytes * a.msgs / dt / 1e9, 'check': chk}))
