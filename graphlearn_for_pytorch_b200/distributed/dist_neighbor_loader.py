"""DistNeighborLoader: distributed neighbour sampling from node seeds (reference
python/distributed/dist_neighbor_loader.py:29-118)."""
from typing import Optional

import torch

from ..sampler import NodeSamplerInput, RemoteSamplerInput, SamplingConfig, SamplingType
from ..typing import Split
from .dist_dataset import DistDataset
from .dist_loader import DistLoader
from .dist_options import AllDistSamplingWorkerOptions


class DistNeighborLoader(DistLoader):
  """Distributed NeighborLoader (node seeds).  `input_nodes`: tensor, (ntype, tensor), or in
  remote mode a Split / path(s) resolved on the server."""

  def __init__(self, data: Optional[DistDataset], num_neighbors, input_nodes, batch_size: int = 1,
               shuffle: bool = False, drop_last: bool = False, with_edge: bool = False,
               with_weight: bool = False, edge_dir: str = 'out', collect_features: bool = False,
               to_device: Optional[torch.device] = None, random_seed: Optional[int] = None,
               worker_options: Optional[AllDistSamplingWorkerOptions] = None):
    if isinstance(input_nodes, tuple):
      input_type, seeds = input_nodes
    else:
      input_type, seeds = None, input_nodes
    from ..sampler import RemoteNodePathSamplerInput, RemoteNodeSplitSamplerInput
    if isinstance(seeds, Split):
      input_data = RemoteNodeSplitSamplerInput(seeds, input_type)
    elif isinstance(seeds, str):
      input_data = RemoteNodePathSamplerInput(seeds, input_type)
    elif isinstance(seeds, list) and seeds and isinstance(seeds[0], str):
      input_data = [RemoteNodePathSamplerInput(p, input_type) for p in seeds]
    elif isinstance(seeds, RemoteSamplerInput) or (isinstance(seeds, list) and seeds and
                                                   isinstance(seeds[0], RemoteSamplerInput)):
      input_data = seeds
    else:
      input_data = NodeSamplerInput(node=torch.as_tensor(seeds), input_type=input_type)
    cfg = SamplingConfig(SamplingType.NODE, num_neighbors, batch_size, shuffle, drop_last, with_edge,
                         collect_features, False, with_weight, edge_dir, random_seed)
    super().__init__(data, input_data, cfg, to_device, worker_options)
