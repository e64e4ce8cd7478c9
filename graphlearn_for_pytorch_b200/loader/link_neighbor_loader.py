"""LinkNeighborLoader (parity: reference python/loader/link_neighbor_loader.py:27-155)."""
from typing import Optional

import torch

from ..data import Dataset
from ..sampler import NegativeSampling, NeighborSampler
from ..typing import InputEdges, NumNeighbors
from .link_loader import LinkLoader


class LinkNeighborLoader(LinkLoader):
  """Mini-batches of seed LINKS with their multi-hop neighbourhoods and (binary or triplet) negative samples;
  batches carry `edge_label_index` / `edge_label` (binary) or `src_index` / `dst_pos_index` / `dst_neg_index`
  (triplet) in local indices (reference: python/loader/link_neighbor_loader.py:27-170)."""
  def __init__(self, data: Dataset, num_neighbors: NumNeighbors,
               neighbor_sampler: Optional[NeighborSampler] = None, edge_label_index: InputEdges = None,
               edge_label: Optional[torch.Tensor] = None, neg_sampling: Optional[NegativeSampling] = None,
               with_edge: bool = False, with_weight: bool = False, batch_size: int = 1,
               shuffle: bool = False, drop_last: bool = False, strategy: str = 'random',
               device: torch.device = None, seed: Optional[int] = None, **kwargs):
    if neighbor_sampler is None:
      neighbor_sampler = NeighborSampler(data.graph, num_neighbors=num_neighbors, strategy=strategy,
                                         with_edge=with_edge, with_weight=with_weight,
                                         with_neg=neg_sampling is not None, device=device,
                                         edge_dir=data.edge_dir, seed=seed)
    super().__init__(data=data, link_sampler=neighbor_sampler, edge_label_index=edge_label_index,
                     edge_label=edge_label, neg_sampling=neg_sampling,
                     device=neighbor_sampler.device if device is None else device,
                     batch_size=batch_size, shuffle=shuffle, drop_last=drop_last, seed=seed, **kwargs)
