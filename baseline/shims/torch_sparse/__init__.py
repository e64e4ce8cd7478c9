"""Minimal stand-in for the `torch_sparse` dependency of the UNMODIFIED reference
(graphlearn_torch/python/utils/topo.py:19,53-69 uses SparseTensor only to sort a COO into
CSR).  torch_sparse is not installable offline; this shim is a *dependency*, not part of
the product and not part of the reference."""
import torch


class _Storage(object):
  def __init__(self, rowptr, col, value):
    self._rowptr, self._col, self._value = rowptr, col, value

  def rowptr(self):
    return self._rowptr

  def col(self):
    return self._col

  def value(self):
    return self._value


class SparseTensor(object):
  def __init__(self, row=None, col=None, value=None, sparse_sizes=None, **kwargs):
    n_rows = int(sparse_sizes[0]) if sparse_sizes is not None else int(row.max()) + 1
    perm = torch.argsort(col, stable=True)
    perm = perm[torch.argsort(row[perm], stable=True)]
    counts = torch.bincount(row, minlength=n_rows)
    rowptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=row.device)
    torch.cumsum(counts, 0, out=rowptr[1:])
    self.storage = _Storage(rowptr, col[perm], value[perm] if value is not None else None)
