"""RandomPartitioner (parity: reference python/partition/random_partitioner.py:28-86)."""
from typing import List, Optional, Tuple

import torch

from ..typing import NodeType
from .base import PartitionerBase
from .partition_book import GLTPartitionBook, PartitionBook


class RandomPartitioner(PartitionerBase):
  """Uniformly random node ownership (balanced: a shuffled id list dealt round-robin)."""

  def _partition_node(self, ntype: Optional[NodeType] = None) -> Tuple[List[torch.Tensor], PartitionBook]:
    n = self._num_nodes(ntype)
    perm = torch.randperm(n)
    pb = torch.empty(n, dtype=torch.int64)
    pb[perm] = torch.arange(n, dtype=torch.int64) % self.num_parts
    ids = [torch.where(pb == p)[0] for p in range(self.num_parts)]
    return ids, GLTPartitionBook(pb)


class RangePartitioner(PartitionerBase):
  """Contiguous id ranges per partition (the layout the NVLink P2P tables use).  Combine
  with a hotness reordering of the ids (data.sort_by_in_degree) to balance load."""

  def _partition_node(self, ntype: Optional[NodeType] = None):
    n = self._num_nodes(ntype)
    # balanced bounds floor(n*p/parts): never inverted, and empty only when n < num_parts (a
    # partition may then own nothing; the tensor book below handles that, RangePartitionBook
    # -- which requires non-empty ranges -- is only built by callers that checked n >= parts)
    bounds = [(n * p) // self.num_parts for p in range(self.num_parts + 1)]
    ids = [torch.arange(bounds[p], bounds[p + 1], dtype=torch.int64) for p in range(self.num_parts)]
    pb = torch.empty(n, dtype=torch.int64)
    for p in range(self.num_parts):
      pb[bounds[p]:bounds[p + 1]] = p
    return ids, GLTPartitionBook(pb)
