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
  """All-gather *views*: returns [view of rank 0's shard, ..., view of rank W-1's shard].

  `t` is copied once into a cudaMalloc'ed PeerBuffer (the symmetric-heap segment of this
  rank); every other rank maps that segment with CUDA IPC **on its own device** with lazy
  peer access, so entry r of the result is a tensor on the *local* device whose storage is
  rank r's HBM: kernels launched here dereference it over NVLink.  The own entry aliases
  the local segment -- drop `t` and keep the returned view to avoid a duplicate.
  """
  rank, world = world_info(group)
  if world == 1:
    return [t]
  assert t.is_cuda and t.is_contiguous()
  nat = require_native()
  my_dev = t.device.index
  nbytes = t.numel() * t.element_size()
  buf = nat.PeerBuffer.allocate(my_dev, nbytes)
  own = buf.as_tensor(t.dtype, list(t.shape))
  own.copy_(t)
  torch.cuda.synchronize(t.device)
  meta = (buf.handle(), nbytes, str(t.dtype), list(t.shape))
  gathered = [None] * world
  dist.all_gather_object(gathered, meta, group=group)
  out = []
  for r, (handle, nb, dtype_s, shape) in enumerate(gathered):
    if r == rank:
      out.append(own)
      continue
    peer = nat.PeerBuffer.open(handle, my_dev, nb)
    out.append(peer.as_tensor(getattr(torch, dtype_s.split('.')[-1]), shape))
  # nobody may free its segment before every peer has mapped it
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
