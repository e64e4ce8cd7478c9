"""Feature-lookup bandwidth (GB/s = numel*bytes / time / 2^30), device-timed.

Counterpart of the reference's benchmarks/api/bench_feature.py:27-61 (`split_ratio` of the rows in
HBM, the rest in pinned host memory).  Sweeps the split ratio and the row dtype.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphlearn_for_pytorch_b200 as glt  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument('--rows', type=int, default=2_449_029)
p.add_argument('--dim', type=int, default=128)
p.add_argument('--ids', type=int, default=400_000, help='rows gathered per lookup (~unique nodes of one batch)')
p.add_argument('--iters', type=int, default=50)
p.add_argument('--impl', default='ours', choices=['ours', 'reference'],
               help="'reference': the unmodified reference (baseline/_ref) Feature on the same box")
args = p.parse_args()
if args.impl == 'reference':
  ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  sys.path.insert(0, os.path.join(ROOT, 'baseline', 'shims'))
  sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))
  try:
    import graphlearn_torch as rglt
    dev = torch.device('cuda', 0)
    results, held = [], []
    for dtype in (torch.float32, torch.bfloat16):
      full = torch.randn(args.rows, args.dim).to(dtype)
      for ratio in (1.0, 0.2, 0.0):
        # a fresh host tensor per Feature: the reference page-locks (cudaHostRegister) the cold part in place and
        # fails on memory that an earlier Feature already registered
        held.append(full.clone())   # kept alive: a freed-and-reused host block would still be registered
        feat = rglt.data.Feature(held[-1], split_ratio=ratio, device_group_list=[rglt.data.DeviceGroup(0, [0])], device=0)
        gen = torch.Generator(device=dev); gen.manual_seed(0)
        ids = [torch.randint(0, args.rows, (args.ids,), device=dev, generator=gen) for _ in range(args.iters + 3)]
        for i in ids[:3]:
          feat[i]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in ids[3:]:
          out = feat[i]
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        results.append({'dtype': str(dtype), 'split_ratio': ratio, 'ms': ms,
                        'GB_per_s': out.numel() * out.element_size() / (ms / 1e3) / 2 ** 30})
        del feat
    print(json.dumps({'impl': 'reference', 'metric': 'feature lookup GB/s', 'rows_per_lookup': args.ids, 'dim': args.dim,
                      'results': results}))
  except Exception as ex:  # noqa: BLE001
    print(json.dumps({'impl': 'reference', 'unavailable': f'{type(ex).__name__}: {str(ex)[:300]}'}))
  sys.exit(0)
dev = torch.device('cuda', 0)
results = []
for dtype in (torch.float32, torch.bfloat16):
  full = torch.randn(args.rows, args.dim).to(dtype)
  for ratio in (1.0, 0.2, 0.0):
    feat = glt.data.Feature(full, split_ratio=ratio, device=0, dtype=dtype)
    gen = torch.Generator(device=dev); gen.manual_seed(0)
    ids = [torch.randint(0, args.rows, (args.ids,), device=dev, generator=gen) for _ in range(args.iters + 3)]
    for i in ids[:3]:
      feat[i]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in ids[3:]:
      out = feat[i]
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    gbs = out.numel() * out.element_size() / (ms / 1e3) / 2 ** 30
    results.append({'dtype': str(dtype), 'split_ratio': ratio, 'ms': ms, 'GB_per_s': gbs})
    del feat
print(json.dumps({'impl': 'ours', 'metric': 'feature lookup GB/s', 'rows_per_lookup': args.ids, 'dim': args.dim,
                  'results': results, 'reference_published_A100_GB_per_s_derived': 11.1}))
