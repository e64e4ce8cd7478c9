"""Sampler inputs / outputs / config (API parity: reference python/sampler/base.py:28-462,
themselves mirrors of torch_geometric.sampler.base)."""
import math
from abc import ABC, abstractmethod
from dataclasses import dataclass
from enum import Enum
from typing import Any, Dict, List, Literal, NamedTuple, Optional, Tuple, Union

import torch

from ..typing import EdgeType, NodeType, NumNeighbors, Split
from ..utils.mixin import CastMixin


class EdgeIndex(NamedTuple):
  """One bipartite layer of the PyG-v1 `adjs` list."""
  edge_index: torch.Tensor
  e_id: Optional[torch.Tensor]
  size: Tuple[int, int]

  def to(self, *args, **kwargs):
    return EdgeIndex(self.edge_index.to(*args, **kwargs),
                     self.e_id.to(*args, **kwargs) if self.e_id is not None else None, self.size)


@dataclass
class NodeSamplerInput(CastMixin):
  """Seed nodes (+ node type for hetero graphs)."""
  node: torch.Tensor
  input_type: Optional[NodeType] = None

  def __getitem__(self, index) -> 'NodeSamplerInput':
    if not isinstance(index, torch.Tensor):
      index = torch.tensor(index, dtype=torch.long)
    return NodeSamplerInput(self.node[index.to(self.node.device)], self.input_type)

  def __len__(self):
    return self.node.numel()

  def share_memory(self):
    self.node.share_memory_()
    return self

  def to(self, device):
    self.node = self.node.to(device)
    return self


class NegativeSamplingMode(Enum):
  binary = 'binary'    # random negative (src, dst) pairs
  triplet = 'triplet'  # random negative dst per positive src


@dataclass
class NegativeSampling(CastMixin):
  mode: NegativeSamplingMode
  amount: Union[int, float] = 1
  weight: Optional[torch.Tensor] = None

  def __init__(self, mode, amount: Union[int, float] = 1, weight: Optional[torch.Tensor] = None):
    self.mode = NegativeSamplingMode(mode)
    self.amount = amount
    self.weight = weight
    if self.amount <= 0:
      raise ValueError(f"'amount' must be positive (got {self.amount})")
    if self.is_triplet():
      if self.amount != math.ceil(self.amount):
        raise ValueError(f"'amount' must be an integer in triplet mode (got {self.amount})")
      self.amount = math.ceil(self.amount)

  def is_binary(self) -> bool:
    return self.mode == NegativeSamplingMode.binary

  def is_triplet(self) -> bool:
    return self.mode == NegativeSamplingMode.triplet

  def share_memory(self):
    if self.weight is not None:
      self.weight.share_memory_()
    return self

  def to(self, device):
    if self.weight is not None:
      self.weight = self.weight.to(device)
    return self


@dataclass
class EdgeSamplerInput(CastMixin):
  """Seed links (+ labels, edge type, negative sampling config)."""
  row: torch.Tensor
  col: torch.Tensor
  label: Optional[torch.Tensor] = None
  input_type: Optional[EdgeType] = None
  neg_sampling: Optional[NegativeSampling] = None

  def __getitem__(self, index) -> 'EdgeSamplerInput':
    if not isinstance(index, torch.Tensor):
      index = torch.tensor(index, dtype=torch.long)
    index = index.to(self.row.device)
    return EdgeSamplerInput(self.row[index], self.col[index],
                            self.label[index] if self.label is not None else None,
                            self.input_type, self.neg_sampling)

  def __len__(self):
    return self.row.numel()

  def share_memory(self):
    self.row.share_memory_()
    self.col.share_memory_()
    if self.label is not None:
      self.label.share_memory_()
    if self.neg_sampling is not None:
      self.neg_sampling.share_memory()
    return self

  def to(self, device):
    self.row = self.row.to(device)
    self.col = self.col.to(device)
    if self.label is not None:
      self.label = self.label.to(device)
    if self.neg_sampling is not None:
      self.neg_sampling.to(device)
    return self


@dataclass
class SamplerOutput(CastMixin):
  """Sampled homogeneous subgraph: `node` (global ids, seeds first, hop-contiguous),
  relabelled `row`/`col`, optional global `edge` ids, per-hop counts."""
  node: torch.Tensor
  row: torch.Tensor
  col: torch.Tensor
  edge: Optional[torch.Tensor] = None
  batch: Optional[torch.Tensor] = None
  num_sampled_nodes: Optional[Union[List[int], torch.Tensor]] = None
  num_sampled_edges: Optional[Union[List[int], torch.Tensor]] = None
  device: Optional[torch.device] = None
  metadata: Optional[Any] = None


@dataclass
class HeteroSamplerOutput(CastMixin):
  node: Dict[NodeType, torch.Tensor]
  row: Dict[EdgeType, torch.Tensor]
  col: Dict[EdgeType, torch.Tensor]
  edge: Optional[Dict[EdgeType, torch.Tensor]] = None
  batch: Optional[Dict[NodeType, torch.Tensor]] = None
  num_sampled_nodes: Optional[Dict[NodeType, Union[List[int], torch.Tensor]]] = None
  num_sampled_edges: Optional[Dict[EdgeType, Union[List[int], torch.Tensor]]] = None
  edge_types: Optional[List[EdgeType]] = None
  input_type: Optional[Union[NodeType, EdgeType]] = None
  device: Optional[torch.device] = None
  metadata: Optional[Any] = None

  def get_edge_index(self):
    edge_index = {k: torch.stack([v, self.col[k]]) for k, v in self.row.items()}
    if self.edge_types is not None:
      for etype in self.edge_types:
        if edge_index.get(etype) is None:
          edge_index[etype] = torch.empty((2, 0), dtype=torch.long, device=self.device)
    return edge_index


@dataclass
class NeighborOutput(CastMixin):
  """One-hop result: flat neighbour ids, per-seed counts, optional edge ids."""
  nbr: torch.Tensor
  nbr_num: torch.Tensor
  edge: Optional[torch.Tensor] = None

  def to(self, device):
    return NeighborOutput(self.nbr.to(device), self.nbr_num.to(device),
                          self.edge.to(device) if self.edge is not None else None)


class SamplingType(Enum):
  NODE = 0
  LINK = 1
  SUBGRAPH = 2
  RANDOM_WALK = 3


@dataclass
class SamplingConfig:
  sampling_type: SamplingType
  num_neighbors: Optional[NumNeighbors]
  batch_size: int
  shuffle: bool
  drop_last: bool
  with_edge: bool
  collect_features: bool
  with_neg: bool
  with_weight: bool = False
  edge_dir: Literal['in', 'out'] = 'out'
  seed: Optional[int] = None


class BaseSampler(ABC):
  """Interface every sampler implements: `sample_from_nodes`, `sample_from_edges`, `subgraph`
  (reference: python/sampler/base.py:355-411)."""
  @abstractmethod
  def sample_from_nodes(self, inputs: NodeSamplerInput, **kwargs):
    ...

  @abstractmethod
  def sample_from_edges(self, inputs: EdgeSamplerInput, **kwargs):
    ...

  @abstractmethod
  def subgraph(self, inputs: NodeSamplerInput) -> SamplerOutput:
    ...


class RemoteSamplerInput(ABC):
  """Sampler input that is resolved on the sampling server."""
  @abstractmethod
  def to_local_sampler_input(self, dataset, **kwargs):
    ...


class RemoteNodePathSamplerInput(RemoteSamplerInput):
  """Seeds stored in a file that the SERVER can read (`torch.load(node_path)`); the client only ships the path."""
  def __init__(self, node_path: str, input_type: Optional[str] = None):
    self.node_path = node_path
    self.input_type = input_type

  def to_local_sampler_input(self, dataset, **kwargs) -> NodeSamplerInput:
    return NodeSamplerInput(node=torch.load(self.node_path), input_type=self.input_type)


class RemoteNodeSplitSamplerInput(RemoteSamplerInput):
  """Seeds = the server-side dataset's train / valid / test split (`typing.Split`)."""
  def __init__(self, split: Split, input_type: Optional[str] = None):
    self.split = split
    self.input_type = input_type

  def to_local_sampler_input(self, dataset, **kwargs) -> NodeSamplerInput:
    idx = {Split.train: dataset.train_idx, Split.valid: dataset.val_idx,
           Split.test: dataset.test_idx}[Split(self.split)]
    if isinstance(idx, dict):
      idx = idx[self.input_type]
    return NodeSamplerInput(node=idx, input_type=self.input_type)
