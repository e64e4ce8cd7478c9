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
    per = (n + self.num_parts - 1) // self.num_parts
    pb = torch.arange(n, dtype=torch.int64) // per
    ids = [torch.arange(p * per, min((p + 1) * per, n)) for p in range(self.num_parts)]
    return ids, GLTPartitionBook(pb)
