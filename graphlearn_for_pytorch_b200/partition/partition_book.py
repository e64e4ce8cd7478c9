"""Partition books (parity: reference python/partition/base.py:30-37, partition_book.py:6-72).

`RangePartitionBook` is the book the NVLink data plane is built around: contiguous id ranges
per partition, so owner(v) is a tiny in-kernel scan (csrc/cuda/device_utils.cuh: load_row /
row_ptr) instead of a tensor lookup + masked_select + RPC routing.
"""
from typing import List, Tuple

import torch


class PartitionBook(object):
  """Maps global ids -> partition ids."""

  def __getitem__(self, indices) -> torch.Tensor:
    raise NotImplementedError

  @property
  def offset(self):
    return None


class GLTPartitionBook(PartitionBook, torch.Tensor):
  """A plain tensor book: book[id] = partition id."""

  def __getitem__(self, indices) -> torch.Tensor:
    # The lookup result is a PLAIN tensor: a subclass result would infect every mask / gather derived from it
    # (each op then detours through Python's __torch_function__) and, worse, the RPC pickler only takes the
    # zero-copy tensor-table path for exact `torch.Tensor`s -- subclass instances are pickled storage by storage.
    return torch.Tensor.__getitem__(self.as_subclass(torch.Tensor), indices)


class RangePartitionBook(PartitionBook):
  """Partition p owns ids in [bounds[p-1], bounds[p]) (bounds are the *end* ids, as in the
  reference); ids are also local-index addressable through `id2index`."""

  def __init__(self, partition_ranges: List[Tuple[int, int]], partition_idx: int):
    if not all(r[0] < r[1] for r in partition_ranges):
      raise ValueError('all partition ranges must be non-empty [start, end)')
    if not all(r[0] == (partition_ranges[i - 1][1] if i > 0 else 0) for i, r in enumerate(partition_ranges)):
      raise ValueError('partition ranges must be contiguous and start at 0')
    self.partition_bounds = torch.tensor([end for _, end in partition_ranges], dtype=torch.long)
    self.partition_idx = partition_idx
    self._id2index = OffsetId2Index(partition_ranges[partition_idx][0])

  def __getitem__(self, indices: torch.Tensor) -> torch.Tensor:
    indices = torch.as_tensor(indices)
    return torch.searchsorted(self.partition_bounds.to(indices.device), indices, right=True)

  @property
  def device(self):
    return self.partition_bounds.device

  @property
  def id2index(self):
    return self._id2index

  def id_filter(self, node_pb: PartitionBook, partition_idx: int) -> torch.Tensor:
    start = int(self.partition_bounds[partition_idx - 1]) if partition_idx > 0 else 0
    end = int(self.partition_bounds[partition_idx])
    return torch.arange(start, end)

  @property
  def bounds(self) -> List[int]:
    """[0, end_0, end_1, ...] as expected by the P2P graph/feature tables."""
    return [0] + self.partition_bounds.tolist()


class OffsetId2Index(object):
  """id -> local row for a contiguous range: id - offset."""

  def __init__(self, offset: int):
    self.offset = offset

  def __getitem__(self, ids: torch.Tensor) -> torch.Tensor:
    local = ids - self.offset
    return local

  def to(self, device):
    return self
