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
    out_dev = row.device
    # the sort is data preparation, not part of any measured region: use the GPU when there is one
    work = torch.device('cuda', torch.cuda.current_device()) if (torch.cuda.is_available() and
                                                                  row.numel() > (1 << 20)) else out_dev
    r, c = row.to(work), col.to(work)
    perm = torch.argsort(c, stable=True)
    perm = perm[torch.argsort(r[perm], stable=True)]
    counts = torch.bincount(r, minlength=n_rows)
    rowptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=work)
    torch.cumsum(counts, 0, out=rowptr[1:])
    v = value.to(work)[perm].to(out_dev) if value is not None else None
    self.storage = _Storage(rowptr.to(out_dev), c[perm].to(out_dev), v)
