"""Channel abstraction between sampling producers and training consumers
(parity: reference python/channel/base.py:20-44)."""
from abc import ABC, abstractmethod
from typing import Dict

import torch

from ..ops import require_native

SampleMessage = Dict[str, torch.Tensor]


def _native_errors():
  nat = require_native()
  return nat.QueueTimeoutError, nat.QueueClosedError


QueueTimeoutError, QueueClosedError = _native_errors()


class ChannelBase(ABC):
  """A FIFO of `SampleMessage`s (dict name -> tensor) between a sampling producer and a consumer: `send`, `recv`,
  `empty` (reference: python/channel/base.py:26-52)."""
  @abstractmethod
  def send(self, msg: SampleMessage, **kwargs):
    ...

  @abstractmethod
  def recv(self, **kwargs) -> SampleMessage:
    ...

  def empty(self) -> bool:
    return False
