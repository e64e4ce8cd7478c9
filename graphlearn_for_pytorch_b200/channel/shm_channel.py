"""Shared-memory channel over the native ring (csrc/cpu/shm_queue.cc, sample_queue.cc).

Parity: reference python/channel/shm_channel.py:24-66 (capacity in messages + byte size,
picklable across spawned processes, optional CUDA pinning, timeout error).
"""
from typing import Optional, Union

import torch

from ..ops import require_native
from ..utils.units import parse_size
from .base import ChannelBase, QueueTimeoutError, SampleMessage


class ShmChannel(ChannelBase):
  """Args:
    capacity: maximum number of in-flight messages.
    shm_size: ring size in bytes or as a string like '256MB'.
  """

  def __init__(self, capacity: int = 128, shm_size: Union[str, int] = '256MB', _attach: Optional[str] = None):
    nat = require_native()
    self.capacity = capacity
    self.shm_size = parse_size(shm_size)
    if _attach is not None:
      self._queue = nat.SampleQueue(_attach)
    else:
      self._queue = nat.SampleQueue(int(capacity), int(self.shm_size))
    self._pinned = False

  @property
  def name(self) -> str:
    return self._queue.name

  def pin_memory(self):
    """Page-lock the whole ring so the consumer's H2D copies run at full PCIe speed."""
    if torch.cuda.is_available() and not self._pinned:
      self._queue.pin_memory()
      self._pinned = True

  def empty(self) -> bool:
    return self._queue.empty()

  def size(self) -> int:
    return self._queue.size()

  def send(self, msg: SampleMessage, **kwargs):
    self._queue.send({k: v for k, v in msg.items() if v is not None})

  def recv(self, timeout_ms: Optional[int] = None, **kwargs) -> SampleMessage:
    return self._queue.recv(int(timeout_ms) if timeout_ms else 0)

  def close(self):
    self._queue.close()

  def __reduce__(self):
    return (_rebuild_shm_channel, (self.capacity, self.shm_size, self.name))


def _rebuild_shm_channel(capacity, shm_size, name):
  return ShmChannel(capacity, shm_size, _attach=name)
