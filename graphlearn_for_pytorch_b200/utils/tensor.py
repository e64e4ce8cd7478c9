"""Tensor helpers (parity: reference python/utils/tensor.py:30-97)."""
from typing import Any, Dict, List, Optional, Union

import numpy as np
import torch


def tensor_equal_with_device(lhs: torch.Tensor, rhs: torch.Tensor) -> bool:
  return lhs.device == rhs.device and torch.equal(lhs, rhs)


def id2idx(ids: Union[List[int], torch.Tensor]) -> torch.Tensor:
  """Dense inverse of an id list: out[ids[i]] = i (size max_id + 1)."""
  ids = ids if isinstance(ids, torch.Tensor) else torch.tensor(ids, dtype=torch.int64)
  ids = ids.to(torch.int64)
  n = int(ids.max().item()) + 1 if ids.numel() > 0 else 0
  out = torch.zeros(n, dtype=torch.int64, device=ids.device)
  out[ids] = torch.arange(ids.numel(), dtype=torch.int64, device=ids.device)
  return out


def convert_to_tensor(data: Any, dtype: Optional[torch.dtype] = None):
  """Recursively turn numpy arrays / lists inside dicts, lists and tuples into tensors."""
  if data is None:
    return None
  if isinstance(data, dict):
    return {k: convert_to_tensor(v, dtype) for k, v in data.items()}
  if isinstance(data, tuple):
    return tuple(convert_to_tensor(v, dtype) for v in data)
  if isinstance(data, list) and len(data) > 0 and not isinstance(data[0], (int, float)):
    return [convert_to_tensor(v, dtype) for v in data]
  if isinstance(data, torch.Tensor):
    return data.to(dtype) if dtype is not None else data
  if isinstance(data, np.ndarray):
    t = torch.from_numpy(data)
    return t.to(dtype) if dtype is not None else t
  return torch.tensor(data, dtype=dtype)


def apply_to_all_tensor(data: Any, tensor_method, *args, **kwargs):
  """Apply `tensor_method(tensor, *args, **kwargs)` to every tensor of a (nested) dict / list / tuple."""
  if data is None:
    return None
  if isinstance(data, dict):
    return {k: apply_to_all_tensor(v, tensor_method, *args, **kwargs) for k, v in data.items()}
  if isinstance(data, (list, tuple)):
    return type(data)(apply_to_all_tensor(v, tensor_method, *args, **kwargs) for v in data)
  if isinstance(data, torch.Tensor):
    return tensor_method(data, *args, **kwargs)
  return data


def share_memory(data: Any):
  """Move CPU tensors (possibly nested) to shared memory in place."""
  def _share(t):
    if t.device.type == 'cpu' and t.numel() > 0:
      t.share_memory_()
    return t
  return apply_to_all_tensor(data, _share)


def squeeze(data: Any):
  return apply_to_all_tensor(data, lambda t: t.squeeze())


def page_lock_in_place(t: torch.Tensor) -> bool:
  """cudaHostRegister the storage behind a (shared-memory) host tensor so kernels can read it where
  it is.  Registers the WHOLE storage (page-aligned mapping base), tolerates "already registered".
  Returns False when the driver refuses -- the caller then falls back to a pinned copy."""
  if t.is_pinned():
    return True
  st = t.untyped_storage()
  if st.nbytes() == 0:
    return False
  err = int(torch.cuda.cudart().cudaHostRegister(st.data_ptr(), st.nbytes(), 0))
  if err not in (0, 712):   # 712 = cudaErrorHostMemoryAlreadyRegistered
    import warnings
    warnings.warn(f'cudaHostRegister failed with error {err}; using a pinned copy instead')
    return False
  return bool(t.is_pinned())
