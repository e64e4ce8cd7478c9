"""NVLink symmetric heap: every rank maps every other rank's shard into its own address
space so kernels can dereference peer HBM directly (ld.global over NVLink 5 / NVSwitch).

Reference counterpart: CUDA-IPC plumbing for the intra-node feature cache only
(csrc/cuda/unified_tensor.cu:168-199,367-381; python/data/unified_tensor.py:92-115), driven
by torch.multiprocessing pickling.  Here it is a collective over the torch.distributed
process group (one process per GPU) and covers graph topology as well as features.
"""
from typing import List, Optional

import torch
import torch.distributed as dist

from ..ops import require_native


def world_info(group=None):
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(group), dist.get_world_size(group)
  return 0, 1


def exchange_peer_tensors(t: torch.Tensor, group=None) -> List[torch.Tensor]:
  """All-gather *views*: returns [tensor of rank 0, ..., tensor of rank W-1] where entry r
  aliases rank r's device memory (CUDA IPC mapping; own entry is `t` itself).

  The caller must keep `t` alive for as long as any peer may read it.
  """
  rank, world = world_info(group)
  if world == 1:
    return [t]
  assert t.is_cuda and t.is_contiguous()
  from torch.multiprocessing.reductions import reduce_tensor
  fn, args = reduce_tensor(t)
  gathered = [None] * world
  dist.all_gather_object(gathered, (fn, args), group=group)
  nat = require_native()
  out = []
  my_dev = t.device.index
  for r, (f, a) in enumerate(gathered):
    if r == rank:
      out.append(t)
      continue
    peer = f(*a)
    nat.enable_peer_access(my_dev, peer.device.index)
    out.append(peer)
  # nobody may free / reuse its shard before every peer has mapped it
  dist.barrier(group=group)
  return out


def exchange_objects(obj, group=None) -> list:
  rank, world = world_info(group)
  if world == 1:
    return [obj]
  out = [None] * world
  dist.all_gather_object(out, obj, group=group)
  return out


def range_bounds(num_rows: int, world: int) -> List[int]:
  """Contiguous, near-equal row ranges: owner(v) is arithmetic / a tiny scan in-kernel."""
  per = (num_rows + world - 1) // world
  return [min(r * per, num_rows) for r in range(world + 1)]
