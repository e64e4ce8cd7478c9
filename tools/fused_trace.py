"""In-situ timeline of the fused layer-1 kernel (GLT_B200_FUSED_TRACE=1).

  GLT_B200_FUSED_TRACE=1 python tools/fused_trace.py [--steps 30] > profiles/fused_trace.txt
  GLT_B200_FUSED_TRACE=1 python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
      tools/fused_trace.py --gpus 8 > profiles/fused_trace_8gpu.txt        # rows resolved over NVLink: rank 0's timeline

Runs the bench engine (pipelined, CUDA graphs, a different seed batch every step, i.e. the real
cache state), then reads the per-CTA clock64 stamps written by the last launch and prints, per tile
slot, the mean/max over CTAs of every event relative to the CTA's start (microseconds).
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('GLT_B200_FUSED_TRACE', '1')

import bench  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--steps', type=int, default=30)
  ap.add_argument('--mhz', type=float, default=1965.0)
  a, rest = ap.parse_known_args()
  sys.argv = [sys.argv[0]] + rest + ['--fused', 'on']
  args = bench.parse_args()
  rank, world, local_rank = bench.setup_dist(args)
  device = torch.device('cuda', local_rank)
  eng, pool = bench.build_ours(args, rank, world, device)
  eng.warmup_and_capture(n_eager=2)
  bs = args.batch
  nb = pool.numel() // bs
  for i in range(a.steps):
    eng.train_step(pool[(i % nb) * bs:(i % nb + 1) * bs].to(device))
  eng.flush()
  torch.cuda.synchronize()
  from graphlearn_for_pytorch_b200.ops import require_native
  tr = require_native().sage_fused_trace().double()
  if world > 1:
    import torch.distributed as dist
    dist.barrier()
  if rank != 0:
    eng.close()
    if world > 1:
      os._exit(0)
    return
  print(f'# fused layer-1 kernel timeline, rank 0 of {world} GPU(s) (remote rows are read from peer HBM over NVLink)')
  start = tr[:, 0:1]
  used = tr[:, 0] > 0
  rel = (tr - start) / a.mhz  # us
  names = ['table published', 'loaders start', 'first row done', 'tile loaded', 'MMA issued',
           'epilogue start', 'epilogue done']
  print(f'CTAs traced: {int(used.sum())}; kernel end: mean {rel[used, 1].mean():.2f} us, max {rel[used, 1].max():.2f} us')
  ntiles = tr[:, 2]
  for nt in sorted(set(ntiles[used].long().tolist())):
    sel = used & (ntiles == nt)
    print(f'-- CTAs with {nt} tiles: {int(sel.sum())}; end mean {rel[sel, 1].mean():.2f} max {rel[sel, 1].max():.2f}')
    for i in range(min(nt, 4)):
      row = []
      for e, n in enumerate(names):
        v = rel[sel, 4 + 7 * i + e]
        row.append(f'{n} {v.mean():.2f}/{v.max():.2f}')
      print(f'   tile {i}: ' + ' | '.join(row))
  eng.close()
  if world > 1:
    sys.stdout.flush()
    os._exit(0)


if __name__ == '__main__':
  main()
