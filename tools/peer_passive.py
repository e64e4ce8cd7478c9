"""Two-rank run in which ONLY rank 0 launches kernels over a graph + feature table partitioned across both GPUs while
rank 1 stays passive (it only owns memory): the setting in which one rank can be profiled with ncu (kernel replay needs
no cooperation from the peer) to read the NVLink counters of the in-kernel peer loads.

  tools/ncu_peer.sh            # launches rank 1 plainly and rank 0 under ncu, writes gpurun_out/nvlink_*.csv
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphlearn_for_pytorch_b200 as glt  # noqa: E402
from graphlearn_for_pytorch_b200.parallel import PartitionedFeature, PartitionedGraph, range_bounds  # noqa: E402
from graphlearn_for_pytorch_b200.utils.synthetic import rmat_csr_shard  # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
local = int(os.environ.get('LOCAL_RANK', rank))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
N, E, F = 2_449_029, 123_718_280, 128
bounds = range_bounds(N, world)
shard = rmat_csr_shard(N, E // 2, bounds[rank], bounds[rank + 1], seed=0, device=dev)
pg = PartitionedGraph(shard, bounds, dev)
local_rows = torch.randn(bounds[rank + 1] - bounds[rank], F, device=dev).to(torch.bfloat16)
pf = PartitionedFeature(local_rows, bounds, dev)          # no hot replica: every remote row crosses NVLink
nat = glt.ops.require_native()
dist.barrier()
if rank == 0:
  fan, bs, H = [15, 10, 5], 1024, 256
  arena = nat.SamplerArena(local, bs, fan, False, N)
  W = (torch.randn(H, 2 * F, device=dev) * 0.05).to(torch.bfloat16)
  wp = nat.pack_weight(W)
  bias = torch.zeros(H, dtype=torch.bfloat16, device=dev)
  cap = sum(arena.cap_rows[:3])
  Z = torch.zeros(cap, H, dtype=torch.bfloat16, device=dev)
  A = torch.zeros(cap, 2 * F, dtype=torch.bfloat16, device=dev)
  torch.cuda.synchronize()
  torch.cuda.cudart().cudaProfilerStart()
  for it in range(3):
    seeds = torch.randint(0, N, (bs,), device=dev)
    arena.sample(pg.graph.graph_handler, seeds, None, 1, it * 8, False, False, False)      # k_sample_hop: peer CSR reads
    nat.sage_fused(pf.table, arena.nodes, None, F, arena.counters, 3, list(arena.ell[:3]), fan, arena.deg, wp, bias, True,
                   Z, A)                                                                    # fused layer: peer feature rows
    ids = torch.randint(bounds[1], N, (400_000,), device=dev)                              # remote-only ids
    pf[ids]                                                                                 # k_gather_vec over NVLink
  torch.cuda.synchronize()
  torch.cuda.cudart().cudaProfilerStop()
  c = arena.counters.cpu().tolist()
  print('rank0 done; last batch nodes', c[1:5], flush=True)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
