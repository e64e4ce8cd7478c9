"""torch.multiprocessing queue channel (parity: reference python/channel/mp_channel.py:21-34)."""
import queue as _queue

import torch.multiprocessing as mp

from .base import ChannelBase, QueueTimeoutError, SampleMessage


class MpChannel(ChannelBase):
  """`torch.multiprocessing.Queue` transport (tensors travel through torch's shared-memory reducers); simple, slower
  than `ShmChannel` (reference: python/channel/mp_channel.py:21-45)."""
  def __init__(self, capacity: int = 128, **kwargs):
    self._q = mp.get_context('spawn').Queue(maxsize=capacity)

  def send(self, msg: SampleMessage, **kwargs):
    self._q.put(msg)

  def recv(self, timeout_ms=None, **kwargs) -> SampleMessage:
    try:
      return self._q.get(timeout=None if not timeout_ms else timeout_ms / 1000.0)
    except _queue.Empty:
      raise QueueTimeoutError('mp channel recv timed out')

  def empty(self) -> bool:
    return self._q.empty()
