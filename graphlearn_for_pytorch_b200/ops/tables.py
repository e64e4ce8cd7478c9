"""IdTable: global id -> dense local id with state kept across hops (the "inducer").

One table per node type gives the hetero inducer.  CPU: native flat hash map
(first-seen order); CUDA: device hash table, ordered for the seed insert,
warp-aggregated (unordered inside a hop, hop-contiguous) afterwards.
Reference counterparts: csrc/cpu/inducer.cc:25-181, csrc/cuda/inducer.cu:75-338.
"""
import threading

import torch

from . import require_native

# Host tables are recycled: a sampler makes one table per batch (per node type), the native table resets in O(1)
# (generation-stamped slots) and grows on demand, so a released table serves any later batch without clearing or
# re-allocating tens of megabytes of slots.
_CPU_POOL, _CPU_POOL_MAX, _CPU_POOL_LOCK = [], 16, threading.Lock()


class IdTable(object):
  def __init__(self, device: torch.device, capacity: int):
    self.device = torch.device(device)
    nat = require_native()
    self.is_cuda = self.device.type == 'cuda'
    self.capacity = max(int(capacity), 16)
    if self.is_cuda:
      self._t = nat.DeviceTable(self.device.index or 0, self.capacity)
    else:
      with _CPU_POOL_LOCK:
        self._t = _CPU_POOL.pop() if _CPU_POOL else None
      if self._t is None:
        self._t = nat.CpuIdTable(self.capacity)
      else:
        self._t.reset()
    self._size = 0

  def __del__(self):
    t = getattr(self, '_t', None)
    if t is not None and not getattr(self, 'is_cuda', True):
      self._t = None
      try:
        with _CPU_POOL_LOCK:
          if len(_CPU_POOL) < _CPU_POOL_MAX:
            _CPU_POOL.append(t)
      except Exception:  # noqa: BLE001  (interpreter shutdown)
        pass

  def reset(self):
    if self.is_cuda:
      self._t.clear()
    else:
      self._t.reset()
    self._size = 0

  def init(self, seeds: torch.Tensor) -> torch.Tensor:
    """Ordered (first-occurrence) insert into an empty table; returns local ids."""
    seeds = seeds.to(self.device, dtype=torch.int64).contiguous()
    if self.is_cuda:
      out = self._t.init_ordered(seeds).to(torch.int64)
      self._size = self._t.size()
    else:
      out = self._t.insert(seeds)
      self._size = self._t.size()
    return out

  def insert(self, keys: torch.Tensor) -> torch.Tensor:
    """Insert (ids < 0 are ignored and map to -1); returns local ids."""
    keys = keys.to(self.device, dtype=torch.int64).contiguous()
    if keys.numel() == 0:
      return keys.clone()
    if self.is_cuda:
      out = self._t.insert(keys).to(torch.int64)
      self._size = self._t.size()
      if self._size > self.capacity:
        raise RuntimeError(f'IdTable overflow: {self._size} > capacity {self.capacity}')
    else:
      out = self._t.insert(keys)
      self._size = self._t.size()
    return out

  def lookup(self, keys: torch.Tensor) -> torch.Tensor:
    keys = keys.to(self.device, dtype=torch.int64).contiguous()
    out = self._t.lookup(keys)
    return out.to(torch.int64)

  def size(self) -> int:
    return self._size

  def keys(self, start: int = 0) -> torch.Tensor:
    if self.is_cuda:
      return self._t.nodes[start:self._size].clone()
    return self._t.keys(start)

  @property
  def native(self):
    return self._t
