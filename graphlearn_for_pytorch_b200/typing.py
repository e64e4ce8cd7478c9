"""Basic graph typing helpers (parity: reference python/typing.py:20-93)."""
from enum import Enum
from typing import Dict, List, NamedTuple, Optional, Tuple, Union

import numpy as np
import torch

NodeType = str
EdgeType = Tuple[str, str, str]
EDGE_TYPE_STR_SPLIT = '__'
REVERSE_PREFIX = 'rev_'


def as_str(type: Union[NodeType, EdgeType]) -> str:  # noqa: A002 - keyword name of the reference API
  """'paper' -> 'paper'; ('a','r','b') -> 'a__r__b'."""
  if isinstance(type, str):
    return type
  if isinstance(type, (list, tuple)) and len(type) == 3:
    return EDGE_TYPE_STR_SPLIT.join(type)
  return ''


def from_str(s: str) -> Union[NodeType, EdgeType]:
  parts = s.split(EDGE_TYPE_STR_SPLIT)
  return tuple(parts) if len(parts) == 3 else s


def reverse_edge_type(etype: EdgeType) -> EdgeType:
  """(src, rel, dst) -> (dst, rev_rel | rel-without-rev_, src); self-relations keep the name."""
  src, rel, dst = etype
  if src != dst:
    rel = rel[len(REVERSE_PREFIX):] if rel.startswith(REVERSE_PREFIX) else REVERSE_PREFIX + rel
  return (dst, rel, src)


TensorDataType = Union[torch.Tensor, np.ndarray]
NodeLabel = Union[TensorDataType, Dict[NodeType, TensorDataType]]
NodeIndex = Union[TensorDataType, Dict[NodeType, TensorDataType]]


class Split(Enum):
  train = 'train'
  valid = 'valid'
  test = 'test'


class GraphPartitionData(NamedTuple):
  """Topology of one partition: (rows, cols) global ids, global edge ids, optional weights."""
  edge_index: Tuple[torch.Tensor, torch.Tensor]
  eids: torch.Tensor
  weights: Optional[torch.Tensor] = None


class FeaturePartitionData(NamedTuple):
  """Feature rows owned by a partition (+ optional hot-cache rows of remote ids)."""
  feats: Optional[torch.Tensor]
  ids: Optional[torch.Tensor]
  cache_feats: Optional[torch.Tensor] = None
  cache_ids: Optional[torch.Tensor] = None


HeteroGraphPartitionData = Dict[EdgeType, GraphPartitionData]
HeteroFeaturePartitionData = Dict[Union[NodeType, EdgeType], FeaturePartitionData]

InputNodes = Union[torch.Tensor, NodeType, Tuple[NodeType, torch.Tensor]]
EdgeIndexTensor = Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]
InputEdges = Union[EdgeIndexTensor, EdgeType, Tuple[EdgeType, EdgeIndexTensor]]
NumNeighbors = Union[List[int], Dict[EdgeType, List[int]]]
