"""Layout conversions without torch_sparse (reference python/utils/topo.py:29-91 needs it).

COO -> CSR/CSC runs in the native counting-sort builder (csrc/cpu/cpu_ops.cc) for
CPU tensors and as a device radix sort (torch.sort on packed keys) for CUDA
tensors.  Rows are always column-sorted: strict negative sampling and the
node2vec walker binary-search them.
"""
from typing import Optional, Tuple

import torch

from ..ops import require_native


def ptr2ind(ptr: torch.Tensor) -> torch.Tensor:
  """CSR row pointer -> per-edge row index."""
  n = ptr.numel() - 1
  counts = ptr[1:] - ptr[:-1]
  return torch.repeat_interleave(torch.arange(n, dtype=ptr.dtype, device=ptr.device), counts)


def ind2ptr(ind: torch.Tensor, size: int) -> torch.Tensor:
  counts = torch.bincount(ind, minlength=size)
  ptr = torch.zeros(size + 1, dtype=torch.int64, device=ind.device)
  torch.cumsum(counts, 0, out=ptr[1:])
  return ptr


def _coo_to_compressed(major: torch.Tensor, minor: torch.Tensor, n_major: Optional[int],
                       edge_ids, edge_weights):
  major = major.to(torch.int64).contiguous()
  minor = minor.to(torch.int64).contiguous()
  if n_major is None:
    n_major = int(major.max().item()) + 1 if major.numel() > 0 else 0
  if major.device.type == 'cpu':
    nat = require_native()
    eids = edge_ids.to(torch.int64).contiguous() if edge_ids is not None else None
    w = edge_weights.contiguous() if edge_weights is not None else None
    ptr, ind, oe, ow = nat.coo_to_csr(major, minor, eids, w, n_major, True)
    return ptr, ind, oe, (ow if edge_weights is not None else None)
  # device path: sort by (major, minor) with two stable sorts
  perm = torch.argsort(minor, stable=True)
  perm = perm[torch.argsort(major[perm], stable=True)]
  ptr = ind2ptr(major, n_major)
  ind = minor[perm]
  oe = edge_ids.to(torch.int64)[perm] if edge_ids is not None else perm
  ow = edge_weights[perm] if edge_weights is not None else None
  return ptr, ind, oe, ow


def coo_to_csr(row, col, edge_id=None, edge_weight=None, node_sizes: Optional[Tuple[int, int]] = None):
  """(row, col) -> (indptr over rows, col indices, edge ids, weights)."""
  n = node_sizes[0] if node_sizes is not None else None
  return _coo_to_compressed(row, col, n, edge_id, edge_weight)


def coo_to_csc(row, col, edge_id=None, edge_weight=None, node_sizes: Optional[Tuple[int, int]] = None):
  """(row, col) -> (row indices, indptr over cols, edge ids, weights)."""
  n = node_sizes[1] if node_sizes is not None else None
  ptr, ind, oe, ow = _coo_to_compressed(col, row, n, edge_id, edge_weight)
  return ind, ptr, oe, ow


def csr_to_coo(indptr, indices):
  return ptr2ind(indptr), indices


def rows_are_sorted(indptr: torch.Tensor, indices: torch.Tensor) -> bool:
  """True when every row of the compressed layout lists its minor indices in ascending order
  (O(E) vectorised check: an inversion is only allowed across a row boundary)."""
  if indices.numel() < 2:
    return True
  inv = indices[1:] < indices[:-1]
  if not bool(inv.any()):
    return True
  starts = torch.zeros(indices.numel() + 1, dtype=torch.bool, device=indices.device)
  starts[indptr.to(torch.int64).clamp(max=indices.numel())] = True   # position i starts a row
  return not bool((inv & ~starts[1:-1]).any())


def sort_csr_columns(indptr, indices, edge_ids=None, edge_weights=None):
  """Make every CSR row column-sorted (used when the user hands in a raw CSR)."""
  row = ptr2ind(indptr)
  ptr, ind, oe, ow = _coo_to_compressed(row, indices, indptr.numel() - 1, edge_ids, edge_weights)
  return ptr, ind, oe, ow
