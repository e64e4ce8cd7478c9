"""Device helpers (parity: reference python/utils/device.py:20-54)."""
import threading
from typing import Optional

import torch

_lock = threading.Lock()
_rr = [0]


def get_available_device(device: Optional[torch.device] = None) -> torch.device:
  if device is not None:
    return torch.device(device)
  if torch.cuda.is_available():
    return torch.device('cuda', torch.cuda.current_device())
  return torch.device('cpu')


def assign_device() -> torch.device:
  """Round-robin over visible GPUs (cpu when none)."""
  if not torch.cuda.is_available():
    return torch.device('cpu')
  with _lock:
    idx = _rr[0] % torch.cuda.device_count()
    _rr[0] += 1
  return torch.device('cuda', idx)


def ensure_device(device: torch.device):
  device = torch.device(device)
  if device.type == 'cuda' and torch.cuda.is_available():
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if torch.cuda.current_device() != idx:
      torch.cuda.set_device(idx)
