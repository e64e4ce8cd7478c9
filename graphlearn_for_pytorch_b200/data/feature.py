"""Feature store with a hot GPU tier and a cold pinned-host tier.

API parity: reference python/data/feature.py:32-283.  Rows [0, split_ratio*N) of the
(hotness-sorted) feature tensor are the hot part: sharded evenly over the GPUs of
the caller's DeviceGroup (each group holds one full replica of the hot rows,
peers inside the group are read over NVLink by the gather kernel); the rest is
pinned host memory read in place.  `id2index` (id -> row after reordering) is
applied on the device inside the gather kernel.
"""
import threading
from multiprocessing.reduction import ForkingPickler
from typing import List, Optional

import torch

from ..ops import _load as _native_module
from ..utils.device import get_available_device
from ..parallel.peer import IpcCudaTensor
from .unified_tensor import UnifiedTensor


class DeviceGroup(object):
  """A set of GPUs with mutual peer access (an NVLink clique / one NVSwitch box)."""

  def __init__(self, group_id: int, device_list: List):
    self.group_id = group_id
    self.device_list = [d.index if isinstance(d, torch.device) else int(d) for d in device_list]

  @property
  def size(self):
    return len(self.device_list)


class Feature(object):
  """Feature store with a hot tier in GPU memory and a cold tier in pinned host memory.

  Rows `[0, split_ratio * N)` of `feature_tensor` (order them by hotness with `sort_by_in_degree`, which also returns
  the `id2index` map) are sharded over the GPUs of the caller's `DeviceGroup`; the rest stays in pinned (or shared,
  page-locked in place) host memory.  `feature[ids]` is ONE gather kernel over all tiers, peer GPUs included.
  `cpu_get(ids)` is the host-side lookup used by RPC callees.  Pickling hands GPU shards to other processes through
  raw CUDA-IPC handles opened on the consumer's device.  (Reference: python/data/feature.py:32-283.)"""
  def __init__(self, feature_tensor: torch.Tensor, id2index: Optional[torch.Tensor] = None,
               split_ratio: float = 0.0, device_group_list: Optional[List[DeviceGroup]] = None,
               device: Optional[int] = None, with_gpu: bool = True,
               dtype: torch.dtype = torch.float32):
    self.feature_tensor = feature_tensor.to(dtype) if feature_tensor is not None and \
        feature_tensor.dtype != dtype else feature_tensor
    # a tensor (id -> row), any sequence convertible to one, or an offset map such as the range partition book's
    # `OffsetId2Index` (row = id - offset; reference partition_book.py:50-64), which is kept as an object
    if id2index is not None and not isinstance(id2index, torch.Tensor) and not hasattr(id2index, 'offset'):
      id2index = torch.as_tensor(id2index, dtype=torch.int64)
    self.id2index = id2index
    self.split_ratio = float(split_ratio)
    self.device_group_list = device_group_list
    self.device = device.index if isinstance(device, torch.device) else device
    self.with_gpu = bool(with_gpu) and torch.cuda.is_available()
    self.dtype = dtype
    self._lock = threading.RLock()
    self._unified: Optional[UnifiedTensor] = None
    self._id2index_dev = None
    self._ipc_handle = None
    self._cuda_parts_by_group = None  # group_id -> list of per-device shards
    self._cpu_part = None
    self._shape = list(feature_tensor.shape) if feature_tensor is not None else None

  # ------------------------------------------------------------------ lookup
  def __getitem__(self, ids: torch.Tensor) -> torch.Tensor:
    if not self.with_gpu:
      return self.cpu_get(ids)
    self.lazy_init()
    dev = torch.device('cuda', self.device)
    ids = ids.to(dev, dtype=torch.int64).contiguous()
    if self._id2index_dev is None and self.id2index is not None:
      ids = self.id2index[ids]        # offset map: plain arithmetic on the device
    out = self._unified._table().gather(ids, self._id2index_dev, 0)
    tail = self._shape[1:]
    return out if len(tail) == 1 else out.view(ids.numel(), *tail)

  def cpu_get(self, ids: torch.Tensor) -> torch.Tensor:
    """Host-side lookup (used by the RPC callee in multi-node mode).  When the host copy of the tensor was
    released (a consumer process that only received the GPU shards + the cold host part over IPC), the rows are
    gathered through the unified table on the device and copied back."""
    ids = ids.to('cpu', dtype=torch.int64)
    if self.feature_tensor is not None:
      out = torch.empty((ids.numel(), *self.feature_tensor.shape[1:]), dtype=self.feature_tensor.dtype)
      return self.cpu_get_into(ids, out)
    if not self.with_gpu:
      raise RuntimeError('feature tensor is not available in this process')
    return self.__getitem__(ids).cpu()

  def cpu_get_into(self, ids: torch.Tensor, out: torch.Tensor, pos: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[pos] (or out[:len(ids)]) = rows of `ids`, one fused native pass over all intra-op threads
    (`cpu_gather_rows`: id2index / offset mapping, bounds checks and the row copies in the same loop; no temporary
    for the gathered rows).  Falls back to torch indexing for tensors the native op does not take."""
    ids = ids.to('cpu', dtype=torch.int64).contiguous().view(-1)
    ft = self.feature_tensor
    i2i = self.id2index
    nat = _native_module()
    if (nat is not None and ft.device.type == 'cpu' and ft.is_contiguous() and out.is_contiguous()
        and out.device.type == 'cpu' and out.dtype == ft.dtype and ft.shape[0] > 0
        and (i2i is None or hasattr(i2i, 'offset') or (i2i.device.type == 'cpu' and i2i.dtype == torch.int64))):
      off = int(i2i.offset) if (i2i is not None and hasattr(i2i, 'offset')) else 0
      table = i2i.contiguous() if isinstance(i2i, torch.Tensor) else None
      nat.cpu_gather_rows(ft, ids, table, off, out, None if pos is None else pos.contiguous())
      return out
    rows = ft[i2i[ids] if i2i is not None else ids]
    if pos is None:
      out[:ids.numel()] = rows
    else:
      out[pos] = rows
    return out

  # ------------------------------------------------------------------ init
  def _check_and_set_device(self):
    if self.device is None:
      self.device = get_available_device().index or 0
    if self.device_group_list is None:
      self.device_group_list = [DeviceGroup(0, [self.device])]
    self._group = None
    for g in self.device_group_list:
      if self.device in g.device_list:
        self._group = g
    if self._group is None:
      self._group = DeviceGroup(len(self.device_group_list), [self.device])

  def lazy_init(self):
    if self._unified is not None or not self.with_gpu:
      return
    with self._lock:
      if self._unified is not None:
        return
      self._check_and_set_device()
      if self._ipc_handle is not None:
        self._init_from_ipc()
      else:
        self._split_and_init()
      if isinstance(self.id2index, torch.Tensor):
        self._id2index_dev = self.id2index.to(torch.device('cuda', self.device), dtype=torch.int64)

  def lazy_init_with_ipc_handle(self):
    """Finish the construction of a Feature received from another process (reference feature.py:242-261):
    here `lazy_init` covers both cases -- a pending IPC handle is opened on first use -- so this is an alias."""
    if self._ipc_handle is not None:
      self.lazy_init()

  def _split_and_init(self):
    n = self.feature_tensor.shape[0]
    hot = int(n * self.split_ratio)
    ut = UnifiedTensor(self.device, self.dtype)
    self._cuda_parts_by_group = {}
    if hot > 0:
      hot_rows = self.feature_tensor[:hot]
      for group in self.device_group_list:
        shards, per = [], (hot + group.size - 1) // group.size
        for i, dev in enumerate(group.device_list):
          part = hot_rows[i * per: min((i + 1) * per, hot)]
          if part.shape[0] > 0:
            shards.append(IpcCudaTensor.from_tensor(part, dev))
        self._cuda_parts_by_group[group.group_id] = shards
      if self._group.group_id not in self._cuda_parts_by_group:
        self._cuda_parts_by_group[self._group.group_id] = [IpcCudaTensor.from_tensor(hot_rows, self.device)]
      for shard in self._cuda_parts_by_group[self._group.group_id]:
        ut.append_shared_tensor(shard.local(self.device))
    if hot < n:
      self._cpu_part = self.feature_tensor[hot:]
      ut.append_cpu_tensor(self._cpu_part)
    self._unified = ut

  # ------------------------------------------------------------------ IPC
  def share_ipc(self):
    """Handle for spawned processes: GPU shards travel as raw CUDA IPC handles that the consumer
    maps on ITS device (parallel/peer.py IpcCudaTensor), the host part as shared memory."""
    with self._lock:
      if self._ipc_handle is not None:
        return self._ipc_handle
      if self.with_gpu:
        self.lazy_init()
        # Share the full matrix first: the cold part is a view of it and becomes shared with it.
        # Nothing temporary may be created here -- while a Process is being spawned the pickler
        # hands file descriptors over by NUMBER, so a shared tensor that is collected before the
        # spawn leaves a recycled fd behind ("unable to resize file" in the child).
        full = self.feature_tensor
        if full is not None and not full.is_shared():
          full.share_memory_()
        cpu = self._cpu_part
        if cpu is not None and not cpu.is_shared():
          self._cpu_part = cpu = cpu.clone().share_memory_()
        if isinstance(self.id2index, torch.Tensor):
          self.id2index = self.id2index.cpu().share_memory_()
        return (self._cuda_parts_by_group, cpu, full, self.id2index, self.split_ratio,
                self.device_group_list, self.with_gpu, self.dtype, self._shape)
      if self.feature_tensor is not None:
        self.feature_tensor.share_memory_()
      if isinstance(self.id2index, torch.Tensor):
        self.id2index = self.id2index.cpu().share_memory_()
      return (None, None, self.feature_tensor, self.id2index, self.split_ratio,
              self.device_group_list, self.with_gpu, self.dtype, self._shape)

  @classmethod
  def from_ipc_handle(cls, ipc_handle):
    (_, _, full, id2index, split_ratio, groups, with_gpu, dtype, shape) = ipc_handle
    f = cls(full, id2index, split_ratio, groups, None, with_gpu, dtype)
    f._ipc_handle = ipc_handle
    f._shape = shape
    return f

  def _init_from_ipc(self):
    parts_by_group, cpu, _, _, _, _, _, _, _ = self._ipc_handle
    ut = UnifiedTensor(self.device, self.dtype)
    shards = (parts_by_group or {}).get(self._group.group_id)
    if shards is None and parts_by_group:
      shards = next(iter(parts_by_group.values()))
    for s in shards or []:
      ut.append_shared_tensor(s.local(self.device))
    if cpu is not None:
      self._cpu_part = cpu
      ut.append_cpu_tensor(cpu)
    self._unified = ut

  # ------------------------------------------------------------------ info
  @property
  def shape(self):
    return self._shape

  def size(self, dim):
    return self._shape[dim]

  @property
  def unified_tensor(self):
    self.lazy_init()
    return self._unified


def rebuild_feature(ipc_handle):
  return Feature.from_ipc_handle(ipc_handle)


def reduce_feature(feature: Feature):
  return (rebuild_feature, (feature.share_ipc(),))


ForkingPickler.register(Feature, reduce_feature)
