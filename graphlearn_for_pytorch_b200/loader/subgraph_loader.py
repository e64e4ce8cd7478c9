"""SubGraphLoader (parity: reference python/loader/subgraph_loader.py:27-100): induced
enclosing subgraphs around seed batches (SEAL-style)."""
from typing import Optional

import torch

from ..utils.tracing import nvtx_range

from ..data import Dataset
from ..sampler import NeighborSampler, NodeSamplerInput
from ..typing import InputNodes, NumNeighbors
from .node_loader import NodeLoader


class SubGraphLoader(NodeLoader):
  """Induced sub-graphs around batches of seed nodes (optionally after k-hop expansion); `mapping` locates the
  seeds inside `node` (reference: python/loader/subgraph_loader.py:25-96)."""
  def __init__(self, data: Dataset, input_nodes: InputNodes, num_neighbors: Optional[NumNeighbors] = None,
               neighbor_sampler: Optional[NeighborSampler] = None, batch_size: int = 1,
               shuffle: bool = False, drop_last: bool = False, with_edge: bool = False,
               strategy: str = 'random', device: torch.device = None, seed: Optional[int] = None,
               **kwargs):
    if neighbor_sampler is None:
      neighbor_sampler = NeighborSampler(data.graph, num_neighbors=num_neighbors, strategy=strategy,
                                         with_edge=with_edge, device=device, edge_dir=data.edge_dir,
                                         seed=seed)
    super().__init__(data=data, node_sampler=neighbor_sampler, input_nodes=input_nodes,
                     device=neighbor_sampler.device if device is None else device,
                     batch_size=batch_size, shuffle=shuffle, drop_last=drop_last, seed=seed, **kwargs)

  def __next__(self):
    seeds = next(self._seeds_iter).to(self.sampler.device)
    with nvtx_range('glt.sample'):
      out = self.sampler.subgraph(NodeSamplerInput(node=seeds, input_type=self._input_type))
    out.batch = seeds
    with nvtx_range('glt.collate'):
      return self._collate_fn(out)
