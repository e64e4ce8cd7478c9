"""LinkLoader (parity: reference python/loader/link_loader.py:35-230)."""
from typing import Optional, Tuple, Union

import torch

from ..utils.tracing import nvtx_range

from ..data import Dataset
from ..sampler import BaseSampler, EdgeSamplerInput, NegativeSampling
from ..typing import InputEdges
from .node_loader import NodeLoader, SeedBatcher


def get_edge_label_index(data: Dataset, edge_label_index: InputEdges):
  """Resolve `edge_label_index` into (edge_type | None, [2, E] tensor); a bare edge
  type or None means "every edge of that relation / of the graph"."""
  edge_type = None
  if edge_label_index is None:
    return None, torch.stack(data.get_graph().topo.to_coo()[:2])
  if isinstance(edge_label_index, tuple) and len(edge_label_index) == 3 and \
      all(isinstance(x, str) for x in edge_label_index):
    edge_type = edge_label_index
    return edge_type, torch.stack(data.get_graph(edge_type).topo.to_coo()[:2])
  if isinstance(edge_label_index, (tuple, list)) and len(edge_label_index) == 2 and \
      isinstance(edge_label_index[0], tuple):
    edge_type, ei = edge_label_index
    if ei is None:
      return edge_type, torch.stack(data.get_graph(edge_type).topo.to_coo()[:2])
    return edge_type, torch.as_tensor(ei) if not isinstance(ei, (tuple, list)) else torch.stack(list(ei))
  if isinstance(edge_label_index, (tuple, list)):
    return None, torch.stack([torch.as_tensor(edge_label_index[0]), torch.as_tensor(edge_label_index[1])])
  return None, edge_label_index


class LinkLoader(NodeLoader):
  """Samples subgraphs around mini-batches of links (+ negatives)."""

  def __init__(self, data: Dataset, link_sampler: BaseSampler, edge_label_index: InputEdges = None,
               edge_label: Optional[torch.Tensor] = None, neg_sampling: Optional[NegativeSampling] = None,
               device: torch.device = None, edge_dir: Optional[str] = None, batch_size: int = 1,
               shuffle: bool = False, drop_last: bool = False, seed: Optional[int] = None, **kwargs):
    """edge_dir: accepted for parity with the reference signature (link_loader.py:108); the direction used when
    batches are assembled is the dataset's (`data.edge_dir`), which is also what the samplers are built with."""
    assert edge_dir is None or edge_dir == data.edge_dir, 'edge_dir must match the dataset edge direction'
    self.edge_dir = data.edge_dir
    self.data = data
    self.sampler = link_sampler
    self.neg_sampling = NegativeSampling.cast(neg_sampling)
    self.device = device if device is not None else getattr(link_sampler, 'device', torch.device('cpu'))
    self.edge_type, ei = get_edge_label_index(data, edge_label_index)
    self.edge_label_index = ei
    self._input_type = self.edge_type
    if self.neg_sampling is not None and self.neg_sampling.is_binary() and edge_label is not None:
      # shift labels by one so that 0 always denotes a negative
      if edge_label.dtype in (torch.int32, torch.int64, torch.int16, torch.uint8):
        edge_label = edge_label + 1
    if self.neg_sampling is not None and self.neg_sampling.is_triplet() and edge_label is not None:
      raise ValueError("'edge_label' must be None in triplet negative sampling mode")
    self.input_data = EdgeSamplerInput(row=ei[0].clone(), col=ei[1].clone(), label=edge_label,
                                       input_type=self.edge_type, neg_sampling=self.neg_sampling)
    self.input_t_label = None
    self._batcher = SeedBatcher(torch.arange(ei.shape[1]), batch_size, shuffle, drop_last, seed)

  def __next__(self):
    idx = next(self._seeds_iter)
    with nvtx_range('glt.sample'):
      out = self.sampler.sample_from_edges(self.input_data[idx])
    with nvtx_range('glt.collate'):
      return self._collate_fn(out)
