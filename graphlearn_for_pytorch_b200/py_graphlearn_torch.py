"""`pywrap` facade: the class / function names of the reference's native module (`py_graphlearn_torch`,
python/py_export_glt.cc:47-222) on top of this package's operators.

The reference's Python layer reaches its C++ core through `from .. import py_graphlearn_torch as pywrap`; code written
against that handle (the reference's own operator tests, custom samplers built from `pywrap.CPURandomSampler` +
`pywrap.CPUInducer`, ...) finds the same names here:

  Graph / GraphMode / SubGraph, CPU|CUDA RandomSampler, CPUWeightedSampler, CPU|CUDA RandomNegativeSampler,
  CPU|CUDA Inducer, CPU|CUDA HeteroInducer, CPU|CUDA SubGraphOp, SampleQueue / QueueTimeoutError,
  UnifiedTensor / SharedTensor, RandomSeedManager, cpu|cuda_stitch_sample_results.

Nothing is re-implemented: every class forwards to the operator the loaders use themselves (`csrc/cpu/cpu_ops.cc`,
`csrc/cuda/{sampling,graph_ops,gather}.cu` through `data.Graph`, `sampler.NeighborSampler`, `ops.IdTable`).
"""
import enum
from typing import Dict, Optional, Tuple

import torch

from . import ops as _ops
from .channel import QueueTimeoutError  # noqa: F401  (same exception type the channels raise)
from .data.graph import Graph as _DataGraph, Topology as _Topology
from .data.unified_tensor import UnifiedTensor as _UnifiedTensor
from .ops.tables import IdTable as _IdTable
from .utils.common import RandomSeedManager as _SeedManager


class GraphMode(enum.Enum):
  DMA = 0          # topology resident in device memory ('CUDA' mode of data.Graph)
  ZERO_COPY = 1    # pinned host memory read in place by the kernels


class Graph(object):
  """CSR store (reference include/graph.h).  `init_cpu_from_csr` / `init_cuda_from_csr` build a `data.Graph`."""

  def __init__(self):
    self.graph: Optional[_DataGraph] = None
    self._mode = None

  def _build(self, indptr, indices, edge_ids, edge_weights, mode, device=None):
    eids = edge_ids if (edge_ids is not None and edge_ids.numel() > 0) else None
    w = edge_weights if (edge_weights is not None and edge_weights.numel() > 0) else None
    topo = _Topology((indptr, indices), edge_ids=eids, edge_weights=w, input_layout='CSR', layout='CSR')
    self.graph = _DataGraph(topo, mode, device)
    self.graph.lazy_init()

  def init_cpu_from_csr(self, indptr, indices, edge_ids=None, edge_weights=None):
    self._mode = None
    self._build(indptr, indices, edge_ids, edge_weights, 'CPU')

  def init_cuda_from_csr(self, indptr, indices, device: int, mode: GraphMode = GraphMode.ZERO_COPY, edge_ids=None,
                         edge_weights=None):
    self._mode = mode
    self._build(indptr, indices, edge_ids, edge_weights, 'CUDA' if mode == GraphMode.DMA else 'ZERO_COPY', device)

  def get_row_count(self) -> int:
    return int(self.graph.row_count)

  def get_col_count(self) -> int:
    return int(self.graph.col_count)

  def get_edge_count(self) -> int:
    return int(self.graph.edge_count)

  def get_mode(self):
    return self._mode


class SubGraph(object):
  def __init__(self, nodes=None, rows=None, cols=None, eids=None):
    self.nodes, self.rows, self.cols, self.eids = nodes, rows, cols, eids


class RandomSeedManager(object):
  """`RandomSeedManager.getInstance().setSeed(seed)` (reference include/common.h:36-65)."""
  _instance = None

  @staticmethod
  def getInstance():
    if RandomSeedManager._instance is None:
      RandomSeedManager._instance = RandomSeedManager()
    return RandomSeedManager._instance

  def setSeed(self, seed: int):
    _SeedManager.set_seed(seed)

  def getSeed(self):
    s = _SeedManager.get_seed()
    return 0 if s is None else s


def _sampler_for(graph: Graph, device: torch.device, with_edge: bool, with_weight: bool):
  from .sampler import NeighborSampler
  return NeighborSampler(graph.graph, None, device=device, with_edge=with_edge, with_weight=with_weight)


class _RandomSampler(object):
  _device_type = 'cpu'
  _weighted = False

  def __init__(self, graph: Graph):
    self._graph = graph
    self._samplers = {}

  def _device(self, ids):
    if self._device_type == 'cpu':
      return torch.device('cpu')
    return ids.device if ids.is_cuda else torch.device('cuda', self._graph.graph.device or 0)

  def _get(self, ids, with_edge):
    key = (with_edge, str(self._device(ids)))
    if key not in self._samplers:
      self._samplers[key] = _sampler_for(self._graph, self._device(ids), with_edge, self._weighted)
    return self._samplers[key]

  def sample(self, ids: torch.Tensor, req_num: int) -> Tuple[torch.Tensor, torch.Tensor]:
    out = self._get(ids, False).sample_one_hop(ids, req_num)
    return out.nbr, out.nbr_num

  def sample_with_edge(self, ids: torch.Tensor, req_num: int):
    out = self._get(ids, True).sample_one_hop(ids, req_num)
    return out.nbr, out.nbr_num, out.edge


class CPURandomSampler(_RandomSampler):
  pass


class CPUWeightedSampler(_RandomSampler):
  _weighted = True


class CUDARandomSampler(_RandomSampler):
  _device_type = 'cuda'

  def cal_nbr_prob(self, k, last_prob, nbr_last_prob, nbr_graph_, cur_prob=None):
    """Probability that a node is reached at the next hop (reference random_sampler.cu CalNbrProb); the result is
    returned and, when `cur_prob` is given, also written into it."""
    nbr_g = nbr_graph_.graph if isinstance(nbr_graph_, Graph) else nbr_graph_
    out = self._get(last_prob, False)._nbr_prob(self._graph.graph, nbr_g, last_prob, nbr_last_prob, int(k))
    if cur_prob is not None:
      cur_prob.copy_(out)
    return out


class _NegativeSampler(object):
  _device_type = 'cpu'

  def __init__(self, graph: Graph):
    from .sampler import RandomNegativeSampler
    self._s = RandomNegativeSampler(graph.graph, mode=self._device_type.upper(), edge_dir='out')

  def sample(self, req_num: int, trials_num: int = 5, padding: bool = False):
    return self._s.sample(req_num, trials_num, padding)


class CPURandomNegativeSampler(_NegativeSampler):
  pass


class CUDARandomNegativeSampler(_NegativeSampler):
  _device_type = 'cuda'


class _Inducer(object):
  """Incremental relabelling: `init_node(seeds)` -> unique seeds, `induce_next(srcs, nbrs, nbrs_num)` -> (nodes new
  at this hop, local row = source of each edge, local col = neighbour of each edge)."""
  _device_type = 'cpu'

  def __init__(self, num_nodes: int):
    self._cap = max(int(num_nodes), 16)
    self._table = None

  def _dev(self, t):
    return torch.device('cpu') if self._device_type == 'cpu' else t.device

  def init_node(self, seed: torch.Tensor) -> torch.Tensor:
    self._table = _IdTable(self._dev(seed), self._cap)
    self._table.init(seed)
    return self._table.keys(0)

  def induce_next(self, srcs, nbrs, nbrs_num):
    before = self._table.size()
    cols = self._table.insert(nbrs)
    rows = torch.repeat_interleave(self._table.lookup(srcs), nbrs_num.to(torch.int64))
    return self._table.keys(before), rows, cols


class CPUInducer(_Inducer):
  pass


class CUDAInducer(_Inducer):
  _device_type = 'cuda'


class _HeteroInducer(object):
  _device_type = 'cpu'

  def __init__(self, num_nodes: Dict[str, int]):
    self._caps = {k: max(int(v), 16) for k, v in num_nodes.items()}
    self._tables: Dict[str, _IdTable] = {}

  def _table(self, ntype, like):
    if ntype not in self._tables:
      dev = torch.device('cpu') if self._device_type == 'cpu' else like.device
      self._tables[ntype] = _IdTable(dev, self._caps.get(ntype, 1 << 16))
    return self._tables[ntype]

  def init_node(self, seed: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    self._tables = {}
    out = {}
    for nt, ids in seed.items():
      t = self._table(nt, ids)
      t.init(ids)
      out[nt] = t.keys(0)
    return out

  def induce_next(self, hetero_nbrs):
    """hetero_nbrs: {(src_type, rel, dst_type): (srcs, nbrs, nbrs_num)} -> ({type: new nodes}, {etype: rows},
    {etype: cols}); all relations of a hop are inserted before any local id is read (reference inducer.cc:138-178)."""
    before = {nt: t.size() for nt, t in self._tables.items()}
    cols = {}
    for et, (srcs, nbrs, _) in hetero_nbrs.items():
      t = self._table(et[2], nbrs)
      before.setdefault(et[2], t.size())
      cols[et] = t.insert(nbrs)
    rows = {et: torch.repeat_interleave(self._table(et[0], srcs).lookup(srcs), num.to(torch.int64))
            for et, (srcs, _, num) in hetero_nbrs.items()}
    nodes = {}
    for nt, t in self._tables.items():
      new = t.keys(before.get(nt, 0))
      if new.numel() > 0:
        nodes[nt] = new
    return nodes, rows, cols


class CPUHeteroInducer(_HeteroInducer):
  pass


class CUDAHeteroInducer(_HeteroInducer):
  _device_type = 'cuda'


class _SubGraphOp(object):
  _device_type = 'cpu'

  def __init__(self, graph: Graph):
    self._graph = graph
    self._samplers = {}

  def node_subgraph(self, srcs: torch.Tensor, with_edge: bool = False) -> SubGraph:
    dev = torch.device('cpu') if self._device_type == 'cpu' else srcs.device
    key = (bool(with_edge), str(dev))
    if key not in self._samplers:
      self._samplers[key] = _sampler_for(self._graph, dev, bool(with_edge), False)
    node, rows, cols, eids, _ = self._samplers[key].node_subgraph(srcs)
    # node_subgraph() hands out the loaders' message-flow orientation (row = neighbour side); the native op of the
    # reference reports the adjacency as stored: rows = source side, cols = neighbour side
    return SubGraph(node, cols, rows, eids)


class CPUSubGraphOp(_SubGraphOp):
  pass


class CUDASubGraphOp(_SubGraphOp):
  _device_type = 'cuda'


class SampleQueue(object):
  """Shared-memory message queue of tensor maps (reference include/sample_queue.h): `send`, `receive(timeout_ms)`,
  `empty`, `pin_memory`; picklable by shared-memory name."""

  def __init__(self, capacity: int, buf_size: int, _native=None):
    self._q = _native if _native is not None else _ops.require_native().SampleQueue(int(capacity), int(buf_size))

  def pin_memory(self):
    return self._q.pin_memory()

  def empty(self) -> bool:
    return self._q.empty()

  def send(self, msg: Dict[str, torch.Tensor]):
    self._q.send({k: (v.cpu() if v.is_cuda else v) for k, v in msg.items()})

  def receive(self, timeout_ms: int = 0):
    return self._q.recv(int(timeout_ms))

  def __reduce__(self):
    return (_rebuild_sample_queue, (self._q.name,))


def _rebuild_sample_queue(name):
  return SampleQueue(0, 0, _native=_ops.require_native().SampleQueue(name))


class SharedTensor(object):
  """A device tensor that other processes can map (reference include/unified_tensor.cuh SharedTensor):
  `share_cuda_ipc()` -> picklable handle, `from_cuda_ipc(handle)`."""

  def __init__(self, tensor: Optional[torch.Tensor] = None):
    from .parallel.peer import IpcCudaTensor
    self._ipc = IpcCudaTensor.from_tensor(tensor, tensor.device.index) if tensor is not None else None

  def share_cuda_ipc(self):
    return self._ipc

  def from_cuda_ipc(self, cuda_ipc):
    self._ipc = cuda_ipc

  def tensor(self, device: Optional[int] = None) -> torch.Tensor:
    return self._ipc.local(device if device is not None else torch.cuda.current_device())


class UnifiedTensor(_UnifiedTensor):
  """`data.UnifiedTensor` under the native module's name; `append_shared_tensor` also takes a `SharedTensor`."""

  def append_shared_tensor(self, shared_tensor, *args, **kwargs):
    if isinstance(shared_tensor, SharedTensor):
      shared_tensor = shared_tensor.tensor(self.device if isinstance(self.device, int) else None)
    return super().append_shared_tensor(shared_tensor, *args, **kwargs)

  def share_cuda_ipc(self):
    """The device-resident parts as `SharedTensor`s (host part not included, as in the reference)."""
    out = []
    for h in self.share_ipc()[0]:
      st = SharedTensor()
      st.from_cuda_ipc(h)
      out.append(st)
    return out


def cpu_stitch_sample_results(ids, idx_list, nbrs_list, nbrs_num_list, eids_list):
  """Merge per-partition one-hop results back into request order -> (nbrs, nbrs_num, eids | None)."""
  nat = _ops.require_native()
  eids_list = list(eids_list or [])
  nbrs, num, eids = nat.cpu_stitch(int(ids.numel()), [t.cpu() for t in idx_list], [t.cpu() for t in nbrs_list],
                                   [t.cpu() for t in nbrs_num_list], [t.cpu() for t in eids_list])
  return nbrs, num, (eids if eids_list else None)


def cuda_stitch_sample_results(ids, idx_list, nbrs_list, nbrs_num_list, eids_list):
  from .distributed.dist_neighbor_sampler import stitch_one_hop
  from .sampler import NeighborOutput
  eids_list = list(eids_list or [])
  parts = [(idx, NeighborOutput(nbrs_list[i], nbrs_num_list[i].to(torch.int64), eids_list[i] if eids_list else None))
           for i, idx in enumerate(idx_list)]
  out = stitch_one_hop(int(ids.numel()), parts, ids.device, bool(eids_list))
  return out.nbr, out.nbr_num, out.edge
