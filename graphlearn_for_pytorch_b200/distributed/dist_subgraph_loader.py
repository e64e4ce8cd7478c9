"""DistSubGraphLoader: induced enclosing sub-graphs across partitions (reference
python/distributed/dist_subgraph_loader.py:28-94)."""
from typing import Optional

import torch

from ..sampler import NodeSamplerInput, SamplingConfig, SamplingType
from .dist_dataset import DistDataset
from .dist_loader import DistLoader
from .dist_options import AllDistSamplingWorkerOptions


class DistSubGraphLoader(DistLoader):
  """Distributed SubGraphLoader (induced enclosing subgraphs)."""

  def __init__(self, data: Optional[DistDataset], input_nodes, num_neighbors=None, batch_size: int = 1,
               shuffle: bool = False, drop_last: bool = False, with_edge: bool = False,
               with_weight: bool = False, edge_dir: str = 'out',
               collect_features: bool = False, to_device: Optional[torch.device] = None,
               random_seed: Optional[int] = None, worker_options: Optional[AllDistSamplingWorkerOptions] = None):
    if isinstance(input_nodes, tuple):
      input_type, seeds = input_nodes
    else:
      input_type, seeds = None, input_nodes
    input_data = NodeSamplerInput(node=torch.as_tensor(seeds), input_type=input_type)
    cfg = SamplingConfig(SamplingType.SUBGRAPH, num_neighbors, batch_size, shuffle, drop_last, with_edge,
                         collect_features, False, with_weight, edge_dir, random_seed)
    super().__init__(data, input_data, cfg, to_device, worker_options)
