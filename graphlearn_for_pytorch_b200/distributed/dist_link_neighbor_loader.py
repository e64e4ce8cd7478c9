"""DistLinkNeighborLoader: distributed neighbour sampling from link seeds, with negative sampling (reference
python/distributed/dist_link_neighbor_loader.py)."""
from typing import Optional

import torch

from ..loader.link_loader import get_edge_label_index
from ..sampler import EdgeSamplerInput, NegativeSampling, SamplingConfig, SamplingType
from .dist_dataset import DistDataset
from .dist_loader import DistLoader
from .dist_options import AllDistSamplingWorkerOptions


class DistLinkNeighborLoader(DistLoader):
  """Distributed LinkNeighborLoader (link seeds + negatives)."""

  def __init__(self, data: Optional[DistDataset], num_neighbors, batch_size: int = 1, edge_label_index=None,
               edge_label: Optional[torch.Tensor] = None, neg_sampling: Optional[NegativeSampling] = None,
               shuffle: bool = False, drop_last: bool = False, with_edge: bool = False,
               with_weight: bool = False, edge_dir: str = 'out', collect_features: bool = False,
               to_device: Optional[torch.device] = None, random_seed: Optional[int] = None,
               worker_options: Optional[AllDistSamplingWorkerOptions] = None):
    edge_type, ei = get_edge_label_index(data, edge_label_index)
    neg_sampling = NegativeSampling.cast(neg_sampling)
    if neg_sampling is not None and neg_sampling.is_binary() and edge_label is not None and \
        edge_label.dtype in (torch.int32, torch.int64):
      edge_label = edge_label + 1
    input_data = EdgeSamplerInput(row=ei[0].clone(), col=ei[1].clone(), label=edge_label, input_type=edge_type,
                                  neg_sampling=neg_sampling)
    cfg = SamplingConfig(SamplingType.LINK, num_neighbors, batch_size, shuffle, drop_last, with_edge,
                         collect_features, neg_sampling is not None, with_weight, edge_dir, random_seed)
    super().__init__(data, input_data, cfg, to_device, worker_options)
