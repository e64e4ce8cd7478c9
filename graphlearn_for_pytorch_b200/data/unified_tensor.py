"""UnifiedTensor: one logical [rows, ...] tensor backed by several physical parts
(local HBM, peer-GPU HBM over NVLink, pinned host memory) with a single-kernel
row gather across all of them.

API parity: reference python/data/unified_tensor.py:23-133 and the native class in
csrc/cuda/unified_tensor.cu:168-381.  The gather kernel is csrc/cuda/gather.cu.
"""
from typing import List, Optional

import torch

from ..ops import require_native
from ..utils.tensor import page_lock_in_place


def _enable_peer(cur: int, other: int):
  if cur == other or other < 0:
    return
  nat = require_native()
  nat.enable_peer_access(int(cur), int(other))


class UnifiedTensor(object):
  """Args:
    current_device: CUDA device index that runs lookups.
    dtype: element type of every part.
  """

  def __init__(self, current_device: int, dtype: torch.dtype = torch.float32):
    self.current_device = int(current_device)
    self.dtype = dtype
    self._parts: List[torch.Tensor] = []
    self._part_devices: List[int] = []
    self._handle = None
    self._cpu_part: Optional[torch.Tensor] = None
    self._ipc_parts = None
    self._ipc_cpu = None
    self._keep_ipc = []

  # ------------------------------------------------------------------ build
  def _table(self):
    if self._handle is None:
      nat = require_native()
      h = nat.RowTableHandle(self.current_device)
      for i, p in enumerate(self._parts):
        h.append(p, bool(self._part_remote[i]) if i < len(self._part_remote) else False)
      self._handle = h
    return self._handle

  def _append(self, t: torch.Tensor, dev: int, remote: bool = False):
    assert t.dtype == self.dtype, f'dtype mismatch: {t.dtype} vs {self.dtype}'
    self._parts.append(t)
    self._part_devices.append(dev)
    if not hasattr(self, '_part_remote'):
      self._part_remote = [False] * (len(self._parts) - 1)
    self._part_remote.append(bool(remote))
    self._handle = None

  def append_shared_tensor(self, shared_tensor: torch.Tensor, remote: bool = False):
    """Append a CUDA tensor (possibly on a peer device / opened from CUDA IPC).  remote=True marks rows that live in
    a PEER GPU's HBM (an IPC mapping looks like local memory to the runtime): kernels that treat local and peer rows
    differently (remote-row staging of the fused engine) rely on it."""
    assert shared_tensor.is_cuda
    _enable_peer(self.current_device, shared_tensor.device.index)
    remote = remote or shared_tensor.device.index != self.current_device
    self._append(shared_tensor.contiguous(), shared_tensor.device.index, remote)

  def append_cpu_tensor(self, cpu_tensor: torch.Tensor):
    """Append a host part; it is page-locked so kernels can read it in place."""
    assert cpu_tensor.device.type == 'cpu'
    t = cpu_tensor.contiguous()
    self._cpu_part = t
    if torch.cuda.is_available():
      if not t.is_pinned():
        # a shared-memory part keeps its mapping (all processes read the same physical rows)
        if not (t.is_shared() and page_lock_in_place(t)):
          t = t.pin_memory()
    self._append(t, -1)

  def init_from(self, tensors: List[torch.Tensor], tensor_devices: List[int]):
    """Place CPU tensors: device index >= 0 -> that GPU's HBM, -1 -> pinned host."""
    assert len(tensors) == len(tensor_devices)
    for t, d in zip(tensors, tensor_devices):
      if t is None or t.numel() == 0:
        continue
      t = t.to(self.dtype)
      if d >= 0:
        self.append_shared_tensor(t.to(torch.device('cuda', d)))
      else:
        self.append_cpu_tensor(t.cpu())

  # ------------------------------------------------------------------ lookup
  def __getitem__(self, ids: torch.Tensor) -> torch.Tensor:
    dev = torch.device('cuda', self.current_device)
    ids = ids.to(dev, dtype=torch.int64).contiguous()
    out = self._table().gather(ids, None, 0)
    tail = self._parts[0].shape[1:] if self._parts else ()
    return out.view(ids.numel(), *tail) if len(tail) != 1 else out

  def gather_into(self, ids, out, n_dev=None, id2index=None):
    self._table().gather_into(ids, id2index, n_dev, out)

  # ------------------------------------------------------------------ info
  @property
  def shape(self):
    rows = sum(p.shape[0] for p in self._parts)
    tail = list(self._parts[0].shape[1:]) if self._parts else []
    return [rows] + tail

  @property
  def device(self):
    return self.current_device

  @property
  def numel(self):
    n = 1
    for s in self.shape:
      n *= s
    return n

  def size(self, dim):
    return self.shape[dim]

  def stride(self, dim):
    return self._parts[0].stride(dim) if self._parts else 0

  @property
  def parts(self):
    return list(zip(self._parts, self._part_devices))

  # ------------------------------------------------------------------ IPC
  def share_ipc(self):
    """(picklable GPU parts, shared-memory host part).

    GPU parts are re-homed once into cudaMalloc'ed IPC segments (parallel/peer.py
    IpcCudaTensor): the consumer process maps them on ITS OWN device, which is what lets a
    kernel on GPU k read a shard that lives on GPU j of another process.
    """
    from ..parallel.peer import IpcCudaTensor
    if self._ipc_parts is None:
      ipc_parts = []
      for i, (p, d) in enumerate(zip(self._parts, self._part_devices)):
        if d < 0:
          continue
        h = IpcCudaTensor.from_tensor(p, d)
        self._parts[i] = h.local(self.current_device)  # drop the duplicate copy
        ipc_parts.append(h)
      self._ipc_parts = ipc_parts
      self._handle = None
    # NB: anything created here must stay referenced by `self`: when a Process is being spawned
    # the pickler passes file descriptors by NUMBER (inheritance), so a temporary shared tensor
    # that is collected before the spawn leaves a dangling / recycled fd behind.
    cpu = self._cpu_part
    if cpu is not None and not cpu.is_shared():
      if self._ipc_cpu is None:
        self._ipc_cpu = cpu.clone().share_memory_()
      cpu = self._ipc_cpu
    return self._ipc_parts, cpu

  def from_ipc_handle(self, cuda_ipc_list, cpu_part):
    for t in cuda_ipc_list:
      if not isinstance(t, torch.Tensor):
        self._keep_ipc.append(t)
        t = t.local(self.current_device)
      self.append_shared_tensor(t)
    if cpu_part is not None:
      self.append_cpu_tensor(cpu_part)

  @classmethod
  def new_from_ipc(cls, ipc_handles, current_device: int, dtype: torch.dtype):
    cuda_ipc_list, cpu_part = ipc_handles
    ut = cls(current_device, dtype)
    ut.from_ipc_handle(cuda_ipc_list, cpu_part)
    return ut
