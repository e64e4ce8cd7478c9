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


# --------------------------------------------------------------------------- picklable HBM
_OWNED = {}  # handle bytes -> IpcCudaTensor owned by this process (a process cannot open its own handle)


class IpcCudaTensor(object):
  """A CUDA tensor that can be handed to *another process running on another GPU*.

  torch's own CUDA-IPC pickling re-opens the allocation in the exporter device's context,
  which kernels of a different GPU cannot dereference.  Here the rows live in a
  cudaMalloc'ed PeerBuffer; the pickle carries the raw cudaIpcMemHandle and the consumer
  maps it **on the device that will read it** with lazy peer access, so the UnifiedTensor
  gather kernel can chase the pointer over NVLink.  Reference counterpart:
  csrc/cuda/unified_tensor.cu:168-199 (ShareCUDAIpc) and :367-381 (InitFrom handles).
  """

  def __init__(self, meta, owner_buf=None, owner_tensor=None):
    self._meta = meta  # (handle, nbytes, dtype_name, shape, owner_device, owner_pid)
    self._buf = owner_buf
    self._tensor = owner_tensor
    self._opened = {}

  @classmethod
  def from_tensor(cls, t: torch.Tensor, device: int) -> 'IpcCudaTensor':
    import os
    nat = require_native()
    device = int(device)
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    buf = nat.PeerBuffer.allocate(device, nbytes)
    own = buf.as_tensor(t.dtype, list(t.shape))
    own.copy_(t)
    torch.cuda.synchronize(device)
    meta = (buf.handle(), nbytes, str(t.dtype).split('.')[-1], list(t.shape), device, os.getpid())
    self = cls(meta, buf, own)
    _OWNED[meta[0]] = self
    return self

  @property
  def owner_device(self) -> int:
    return self._meta[4]

  @property
  def shape(self):
    return self._meta[3]

  def local(self, device: int) -> torch.Tensor:
    """A tensor usable by kernels launched on `device`."""
    import os
    device = int(device)
    if self._tensor is not None:  # owner process: plain cudaMalloc memory + peer access
      if device != self.owner_device:
        require_native().enable_peer_access(device, self.owner_device)
      return self._tensor
    handle, nbytes, dtype_s, shape, _, owner_pid = self._meta
    if owner_pid == os.getpid():
      owner = _OWNED.get(handle)
      if owner is None:
        raise RuntimeError('IPC handle exported by this process but its owner was released')
      return owner.local(device)
    if device not in self._opened:
      nat = require_native()
      buf = nat.PeerBuffer.open(handle, device, nbytes)
      self._opened[device] = (buf, buf.as_tensor(getattr(torch, dtype_s), shape))
    return self._opened[device][1]

  def __reduce__(self):
    return (IpcCudaTensor, (self._meta,))
