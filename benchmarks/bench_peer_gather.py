"""Remote feature gather over NVLink: GB/s per GPU when the rows live in PEER HBM and are read
in-kernel by `k_gather_vec` (BASELINE.json: "remote feature-gather GB/s vs 900 GB/s").

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      benchmarks/bench_peer_gather.py [--rows 2449029 --dim 128 --ids 400000]

Every rank owns a contiguous row range (PartitionedFeature, no hot replica) and gathers
  remote : ids drawn only from the OTHER ranks' ranges  (pure NVLink traffic)
  uniform: ids uniform over all rows                    ((W-1)/W remote)
  local  : ids from its own range                       (HBM reference point)
Timed with CUDA events after warm-up, MAX over ranks; a fresh id set every iteration (inputs
larger than L2 anyway: rows*dim*2 B per rank).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphlearn_for_pytorch_b200 as glt  # noqa: E402,F401
from graphlearn_for_pytorch_b200.parallel import PartitionedFeature, range_bounds  # noqa: E402


def main():
  p = argparse.ArgumentParser()
  p.add_argument('--rows', type=int, default=2_449_029)
  p.add_argument('--dim', type=int, default=128)
  p.add_argument('--ids', type=int, default=400_000)
  p.add_argument('--iters', type=int, default=30)
  a = p.parse_args()
  rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
  local = int(os.environ.get('LOCAL_RANK', rank))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
  dist.init_process_group('nccl', device_id=dev)
  bounds = range_bounds(a.rows, world)
  mine = torch.randn(bounds[rank + 1] - bounds[rank], a.dim, device=dev).to(torch.bfloat16)
  pf = PartitionedFeature(mine, bounds, dev)
  table = pf.unified._table()
  g = torch.Generator(device=dev).manual_seed(100 + rank)
  row_bytes = a.dim * 2

  def ids_for(kind):
    if kind == 'uniform' or world == 1:
      return torch.randint(0, a.rows, (a.ids,), device=dev, generator=g)
    lo, hi = bounds[rank], bounds[rank + 1]
    if kind == 'local':
      return torch.randint(lo, hi, (a.ids,), device=dev, generator=g)
    r = torch.randint(0, a.rows - (hi - lo), (a.ids,), device=dev, generator=g)
    return torch.where(r >= lo, r + (hi - lo), r)        # skip the own range

  out = torch.empty(a.ids, a.dim, dtype=torch.bfloat16, device=dev)
  results = {}
  for kind in ('local', 'uniform', 'remote'):
    id_sets = [ids_for(kind) for _ in range(a.iters + 3)]
    for i in range(3):
      table.gather_into(id_sets[i], None, None, out)
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(a.iters):
      table.gather_into(id_sets[3 + i], None, None, out)
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e) / a.iters], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    remote_frac = {'local': 0.0, 'uniform': (world - 1) / world, 'remote': 1.0 if world > 1 else 0.0}[kind]
    gbs = a.ids * row_bytes / (float(ms) * 1e-3) / 1e9
    results[kind] = {'ms': round(float(ms), 4), 'GB_per_s_per_gpu': round(gbs, 1),
                     'remote_fraction': remote_frac,
                     'nvlink_GB_per_s_per_gpu': round(gbs * remote_frac, 1),
                     'fraction_of_900_GB_per_s': round(gbs * remote_frac / 900.0, 3)}
  # check correctness of one remote gather against the owner's rows
  ids = ids_for('remote')[:1000]
  got = table.gather(ids, None, 0)
  owners = torch.bucketize(ids, torch.tensor(bounds[1:], device=dev), right=True)
  ok = True
  for r in range(world):
    m = owners == r
    if m.any():
      ok &= bool(torch.equal(got[m], pf.peers[r][ids[m] - bounds[r]]))
  if rank == 0:
    print(json.dumps({'metric': 'peer-HBM feature gather (in-kernel NVLink loads)', 'n_gpus': world, 'rows': a.rows,
                      'dim': a.dim, 'dtype': 'bf16', 'ids_per_lookup': a.ids, 'verified': ok,
                      'timing': 'CUDA events, max over ranks, fresh ids every iteration', 'results': results}),
          flush=True)
  dist.barrier()
  torch.cuda.synchronize()
  os._exit(0)


if __name__ == '__main__':
  main()
