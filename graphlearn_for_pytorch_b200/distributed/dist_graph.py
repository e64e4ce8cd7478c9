"""DistGraph: local topology + partition books (parity: reference python/distributed/dist_graph.py:28-124)."""
from typing import Dict, Optional, Union

import torch

from ..data import Graph
from ..partition import PartitionBook
from ..typing import EdgeType, NodeType


class DistGraph(object):
  """Args:
    num_partitions / partition_idx: layout of the partitioned dataset.
    local_graph: `Graph` (homo) or Dict[EdgeType, Graph] held by this process.
    node_pb / edge_pb: partition books (tensor books or RangePartitionBook), dicts for hetero.
  """

  def __init__(self, num_partitions: int, partition_idx: int,
               local_graph: Union[Graph, Dict[EdgeType, Graph]], node_pb, edge_pb=None):
    self.num_partitions = num_partitions
    self.partition_idx = partition_idx
    self.local_graph = local_graph
    self.data_cls = 'hetero' if isinstance(local_graph, dict) else 'homo'
    if self.data_cls == 'hetero':
      self.node_types, self.edge_types = [], []
      for et in local_graph.keys():
        self.edge_types.append(et)
        for t in (et[0], et[-1]):
          if t not in self.node_types:
            self.node_types.append(t)
    else:
      self.node_types = self.edge_types = None
    self.node_pb = node_pb
    self.edge_pb = edge_pb

  def lazy_init(self):
    if isinstance(self.local_graph, dict):
      for g in self.local_graph.values():
        g.lazy_init()
    else:
      self.local_graph.lazy_init()

  def get_local_graph(self, etype: Optional[EdgeType] = None) -> Graph:
    if self.data_cls == 'hetero':
      assert etype is not None
      return self.local_graph[etype]
    return self.local_graph

  def _book(self, pb, t):
    if isinstance(pb, dict):
      assert t is not None
      return pb[t]
    return pb

  def get_node_partitions(self, ids: torch.Tensor, ntype: Optional[NodeType] = None) -> torch.Tensor:
    pb = self._book(self.node_pb, ntype)
    dev = getattr(pb, 'device', torch.device('cpu'))
    return pb[ids.to(dev)]

  def get_edge_partitions(self, eids: torch.Tensor, etype: Optional[EdgeType] = None) -> torch.Tensor:
    pb = self._book(self.edge_pb, etype)
    dev = getattr(pb, 'device', torch.device('cpu'))
    return pb[eids.to(dev)]
