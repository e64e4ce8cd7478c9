"""NodeLoader: seed batching + sampler + feature/label collation.

Parity: reference python/loader/node_loader.py:27-115.  Seed batching does not go
through torch.utils.data.DataLoader (pure index slicing; its worker machinery adds
latency for nothing here) and the epoch position is checkpointable.
"""
from typing import Optional, Union

import torch

from ..data import Dataset
from ..sampler import BaseSampler, HeteroSamplerOutput, NodeSamplerInput, SamplerOutput
from ..typing import InputNodes
from .transform import to_data, to_hetero_data


class SeedBatcher(object):
  """Deterministic (seeded) shuffling batcher over an index tensor."""

  def __init__(self, seeds: torch.Tensor, batch_size: int = 1, shuffle: bool = False,
               drop_last: bool = False, seed: Optional[int] = None):
    self.seeds = seeds
    self.batch_size = int(batch_size)
    self.shuffle = shuffle
    self.drop_last = drop_last
    self.seed = seed if seed is not None else int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    self.epoch = 0
    self.pos = 0
    self._order = None
    self._resumed = False

  def __len__(self):
    n = self.seeds.shape[0]
    return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

  def _make_order(self, epoch_index: int):
    if not self.shuffle:
      return None
    g = torch.Generator()
    g.manual_seed(self.seed + epoch_index)
    return torch.randperm(self.seeds.shape[0], generator=g)

  def __iter__(self):
    if self._resumed:
      # load_state_dict() put us in the middle of epoch `self.epoch - 1`: keep the position and the
      # permutation of THAT epoch so the resumed stream continues exactly where it stopped
      self._resumed = False
      return self
    self._order = self._make_order(self.epoch)
    self.pos = 0
    self.epoch += 1
    return self

  def __next__(self):
    n = self.seeds.shape[0]
    if self.pos >= n or (self.drop_last and self.pos + self.batch_size > n):
      raise StopIteration
    idx = slice(self.pos, min(self.pos + self.batch_size, n))
    self.pos += self.batch_size
    if self._order is not None:
      return self.seeds[self._order[idx]]
    return self.seeds[idx]

  def state_dict(self):
    return {'seed': self.seed, 'epoch': self.epoch, 'pos': self.pos}

  def load_state_dict(self, s):
    """Restore a mid-epoch position: the permutation of the interrupted epoch (`epoch - 1`, since
    `epoch` counts started epochs) is recomputed from the seed, and the next `iter()` continues
    it instead of starting a new epoch."""
    self.seed, self.epoch, self.pos = s['seed'], s['epoch'], s['pos']
    if self.epoch > 0:
      self._order = self._make_order(self.epoch - 1)
      self._resumed = True
    else:
      self._order, self._resumed = None, False


class NodeLoader(object):
  """Base of the node-seeded loaders: batches seed ids (`SeedBatcher`: shuffle, drop_last, resumable position),
  turns sampler output into PyG-shaped `Data` / `HeteroData` with features and labels attached
  (reference: python/loader/node_loader.py:27-112)."""
  def __init__(self, data: Dataset, node_sampler: BaseSampler, input_nodes: InputNodes,
               device: torch.device = None, batch_size: int = 1, shuffle: bool = False,
               drop_last: bool = False, seed: Optional[int] = None, **kwargs):
    self.data = data
    self.sampler = node_sampler
    self.input_nodes = input_nodes
    self.device = device if device is not None else getattr(node_sampler, 'device', torch.device('cpu'))
    if isinstance(input_nodes, tuple):
      input_type, input_seeds = input_nodes
    else:
      input_type, input_seeds = None, input_nodes
    if isinstance(input_seeds, str):  # a bare node type: every node of that type
      input_type = input_seeds
      input_seeds = torch.arange(data.get_node_label(input_type).shape[0])
    self._input_type = input_type
    label = self.data.get_node_label(self._input_type)
    self.input_t_label = label.to(self.device) if isinstance(label, torch.Tensor) else None
    self._batcher = SeedBatcher(torch.as_tensor(input_seeds), batch_size, shuffle, drop_last, seed)

  def __len__(self):
    return len(self._batcher)

  def __iter__(self):
    self._seeds_iter = iter(self._batcher)
    return self

  def __next__(self):
    raise NotImplementedError

  def state_dict(self):
    s = {'batcher': self._batcher.state_dict()}
    if hasattr(self.sampler, 'state_dict'):
      s['sampler'] = self.sampler.state_dict()
    return s

  def load_state_dict(self, s):
    self._batcher.load_state_dict(s['batcher'])
    if 'sampler' in s and hasattr(self.sampler, 'load_state_dict'):
      self.sampler.load_state_dict(s['sampler'])

  def _collate_fn(self, sampler_out: Union[SamplerOutput, HeteroSamplerOutput]):
    if isinstance(sampler_out, SamplerOutput):
      x = self.data.node_features[sampler_out.node] if self.data.node_features is not None else None
      y = self.input_t_label[sampler_out.node.to(self.input_t_label.device)] \
        if self.input_t_label is not None else None
      edge_attr = None
      if self.data.edge_features is not None and sampler_out.edge is not None:
        edge_attr = self.data.edge_features[sampler_out.edge]
      return to_data(sampler_out, batch_labels=y, node_feats=x, edge_feats=edge_attr)
    x_dict = {}
    for ntype, ids in sampler_out.node.items():
      feat = self.data.get_node_feature(ntype)
      if feat is not None:
        x_dict[ntype] = feat[ids]
    y_dict = None
    if self.input_t_label is not None and self._input_type in sampler_out.node:
      ids = sampler_out.node[self._input_type]
      y_dict = {self._input_type: self.input_t_label[ids.to(self.input_t_label.device)]}
    edge_attr_dict = {}
    if sampler_out.edge is not None:
      for etype, eids in sampler_out.edge.items():
        efeat = self.data.get_edge_feature(etype)
        if efeat is None:
          from ..typing import reverse_edge_type
          efeat = self.data.get_edge_feature(reverse_edge_type(etype))
        if efeat is not None:
          edge_attr_dict[etype] = efeat[eids]
    return to_hetero_data(sampler_out, batch_label_dict=y_dict, node_feat_dict=x_dict,
                          edge_feat_dict=edge_attr_dict, edge_dir=self.data.edge_dir)


def _loader_repr(self) -> str:
  return f'{self.__class__.__name__}()'


NodeLoader.__repr__ = _loader_repr     # `LinkNeighborLoader()` etc., as in the reference
# the reference keeps its batching torch DataLoader in `_seed_loader` (len() = batches per epoch); same handle here
NodeLoader._seed_loader = property(lambda self: self._batcher)
