"""NeighborLoader (parity: reference python/loader/neighbor_loader.py:27-112)."""
from typing import Optional

import torch

from ..utils.tracing import nvtx_range

from ..data import Dataset
from ..sampler import NeighborSampler, NodeSamplerInput
from ..typing import InputNodes, NumNeighbors
from .node_loader import NodeLoader


class NeighborLoader(NodeLoader):
  """Mini-batch loader with multi-hop neighbour sampling.

  Yields `Data` / `HeteroData` (x, y, edge_index, node, edge, batch, batch_size,
  num_sampled_nodes/edges), or the legacy `(batch_size, n_id, adjs)` triple when
  `as_pyg_v1=True`.
  """

  def __init__(self, data: Dataset, num_neighbors: NumNeighbors, input_nodes: InputNodes,
               neighbor_sampler: Optional[NeighborSampler] = None, batch_size: int = 1,
               shuffle: bool = False, drop_last: bool = False, with_edge: bool = False,
               with_weight: bool = False, strategy: str = 'random',
               device: torch.device = None, as_pyg_v1: bool = False, seed: Optional[int] = None,
               **kwargs):
    if neighbor_sampler is None:
      neighbor_sampler = NeighborSampler(data.graph, num_neighbors=num_neighbors, strategy=strategy,
                                         with_edge=with_edge, with_weight=with_weight, device=device,
                                         edge_dir=data.edge_dir, seed=seed)
    self.as_pyg_v1 = as_pyg_v1
    self.edge_dir = data.edge_dir
    super().__init__(data=data, node_sampler=neighbor_sampler, input_nodes=input_nodes,
                     device=neighbor_sampler.device if device is None else device,
                     batch_size=batch_size, shuffle=shuffle, drop_last=drop_last, seed=seed, **kwargs)

  def __next__(self):
    seeds = next(self._seeds_iter).to(self.sampler.device)
    if not self.as_pyg_v1:
      with nvtx_range('glt.sample'):
        out = self.sampler.sample_from_nodes(NodeSamplerInput(node=seeds, input_type=self._input_type))
      with nvtx_range('glt.collate'):
        return self._collate_fn(out)
    with nvtx_range('glt.sample'):
      return self.sampler.sample_pyg_v1(seeds)
