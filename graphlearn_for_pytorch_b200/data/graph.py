"""Topology (CSR/CSC container) and Graph (device placement of a topology).

API parity: reference python/data/graph.py:28-306.  Differences by design:
  * no torch_sparse: layout conversion uses the native counting-sort builder;
  * 'CUDA' mode keeps column ids as int32 in HBM when they fit (half the bytes
    per sampled neighbour), 'ZERO_COPY' pins the host arrays and lets kernels
    read them in place (cold tier), 'CPU' uses the native C++ samplers;
  * edge weights are uploaded too (the reference never puts them on the GPU,
    include/graph.h:60-71), enabling weighted sampling on device;
  * a Graph can also be assembled from several shards living on different GPUs
    (`Graph.from_shards`) -- the kernels then read peer HBM over NVLink.
"""
import threading
from multiprocessing.reduction import ForkingPickler
from typing import List, Literal, Optional, Tuple, Union

import torch

from ..utils.tensor import page_lock_in_place

from ..ops import require_native
from ..typing import TensorDataType
from ..utils.tensor import convert_to_tensor, share_memory
from ..utils.topo import coo_to_csc, coo_to_csr, ptr2ind, rows_are_sorted, sort_csr_columns


class Topology(object):
  """Graph topology in CSR (edge_dir='out') or CSC (edge_dir='in') layout.

  Args:
    edge_index: [2, E] COO, or (indptr, indices) for 'CSR', or (indices, indptr) for 'CSC'.
    edge_ids: optional global edge ids (default arange(E)).
    edge_weights: optional float weights.
    input_layout: 'COO' | 'CSR' | 'CSC'.
    layout: target layout 'CSR' | 'CSC'.
  """

  def __init__(self, edge_index, edge_ids: Optional[TensorDataType] = None,
               edge_weights: Optional[TensorDataType] = None, input_layout: str = 'COO',
               layout: Literal['CSR', 'CSC'] = 'CSR', num_nodes: Optional[int] = None):
    edge_index = convert_to_tensor(edge_index, dtype=torch.int64)
    row, col = edge_index[0], edge_index[1]
    input_layout = str(input_layout).upper()
    layout = str(layout).upper()
    if input_layout == 'COO':
      assert row.numel() == col.numel()
      num_edges = row.numel()
    elif input_layout == 'CSR':
      num_edges = col.numel()
    elif input_layout == 'CSC':
      num_edges = row.numel()
    else:
      raise RuntimeError(f"'{self.__class__.__name__}': invalid edge layout {input_layout}")
    edge_ids = convert_to_tensor(edge_ids, dtype=torch.int64)
    if edge_ids is None:
      edge_ids = torch.arange(num_edges, dtype=torch.int64, device=row.device)
    assert edge_ids.numel() == num_edges
    edge_weights = convert_to_tensor(edge_weights, dtype=torch.float32)
    if edge_weights is not None:
      assert edge_weights.numel() == num_edges
    self._layout = layout

    if input_layout == layout:
      if layout == 'CSC':
        ind, ptr = row.contiguous(), col.contiguous()
      else:
        ptr, ind = row.contiguous(), col.contiguous()
      # strict negative sampling, node2vec walks and induced sub-graphs binary-search inside a row
      # (csrc/cuda/graph_ops.cu edge_exists, cpu_negative_sample(sorted=True)); a user-supplied
      # compressed layout is therefore brought to the same sorted-minor-index form that the COO
      # path produces (edge ids / weights are permuted along).
      if ind.numel() > 1 and not rows_are_sorted(ptr, ind):
        ptr, ind, edge_ids, edge_weights = sort_csr_columns(ptr, ind, edge_ids, edge_weights)
      self._indptr, self._indices = ptr, ind
      self._edge_ids, self._edge_weights = edge_ids, edge_weights
      return
    if input_layout == 'CSC':
      col = ptr2ind(col)
    elif input_layout == 'CSR':
      row = ptr2ind(row)
    sizes = (num_nodes, num_nodes) if num_nodes is not None else None
    if layout == 'CSR':
      self._indptr, self._indices, self._edge_ids, self._edge_weights = \
        coo_to_csr(row, col, edge_ids, edge_weights, node_sizes=sizes)
    else:
      self._indices, self._indptr, self._edge_ids, self._edge_weights = \
        coo_to_csc(row, col, edge_ids, edge_weights, node_sizes=sizes)

  def to_coo(self):
    """(row, col, edge_ids, weights)."""
    major = ptr2ind(self._indptr)
    if self._layout == 'CSR':
      return major, self._indices, self._edge_ids, self._edge_weights
    return self._indices, major, self._edge_ids, self._edge_weights

  def to_csc(self):
    """(row indices, col indptr, edge_ids, weights)."""
    if self._layout == 'CSC':
      return self._indices, self._indptr, self._edge_ids, self._edge_weights
    row, col, eid, w = self.to_coo()
    return coo_to_csc(row, col, eid, w)

  def to_csr(self):
    """(row indptr, col indices, edge_ids, weights)."""
    if self._layout == 'CSR':
      return self._indptr, self._indices, self._edge_ids, self._edge_weights
    row, col, eid, w = self.to_coo()
    return coo_to_csr(row, col, eid, w)

  @property
  def layout(self):
    return self._layout

  @property
  def indptr(self):
    return self._indptr

  @property
  def indices(self):
    return self._indices

  @property
  def edge_ids(self):
    return self._edge_ids

  @property
  def edge_weights(self):
    return self._edge_weights

  @property
  def degrees(self):
    return self._indptr[1:] - self._indptr[:-1]

  @property
  def row_count(self):
    return self._indptr.shape[0] - 1

  @property
  def edge_count(self):
    return self._indices.shape[0]

  def share_memory_(self):
    share_memory([self._indptr, self._indices, self._edge_ids, self._edge_weights])
    return self

  def __getitem__(self, key):
    return getattr(self, key, None)

  def __setitem__(self, key, value):
    setattr(self, key, value)


def _page_lock(t: torch.Tensor) -> torch.Tensor:
  """Make a host tensor readable by kernels in place.  A shared-memory tensor (a Graph that
  travelled to a spawned process) is registered where it is, so N processes keep ONE copy."""
  if t.is_pinned():
    return t
  if t.is_shared() and t.numel() > 0 and page_lock_in_place(t):
    return t
  return t.pin_memory()


class Graph(object):
  """A topology placed for sampling.

  mode:
    'CPU'        native C++ samplers on host memory
    'ZERO_COPY'  host arrays pinned, kernels dereference them in place (PCIe)
    'CUDA'       arrays resident in HBM (int32 column ids when they fit)
  """

  def __init__(self, topo: Optional[Topology], mode: str = 'ZERO_COPY',
               device: Optional[int] = None):
    self.topo = topo
    self.mode = str(mode).upper()
    assert self.mode in ('CPU', 'ZERO_COPY', 'CUDA')
    if self.mode != 'CPU' and not torch.cuda.is_available():
      self.mode = 'CPU'  # no GPU in this process: host samplers
    self.device = device
    if self.mode != 'CPU' and self.device is None:
      self.device = torch.cuda.current_device()
    if isinstance(self.device, torch.device):
      self.device = self.device.index if self.device.index is not None else 0
    self._handle = None
    self._shards = None
    self._keep = []
    self._col_count = None
    self._lock = threading.RLock()

  # ------------------------------------------------------------------ build
  @classmethod
  def from_shards(cls, shards: List[dict], device: int):
    """Assemble a multi-shard device graph.  Each shard dict holds
    indptr/indices/(eids)/(weights) tensors (any CUDA device or pinned host)
    and the [row_begin, row_end) range it owns."""
    g = cls(None, 'CUDA', device)
    g._shards = shards
    g.lazy_init()
    return g

  def lazy_init(self):
    if self._handle is not None or self.mode == 'CPU':
      return
    with self._lock:
      if self._handle is not None:
        return
      nat = require_native()
      h = nat.GraphHandle(int(self.device))
      if self._shards is not None:
        for s in self._shards:
          h.add_shard(s['indptr'], s['indices'], s.get('eids'), s.get('weights'),
                      int(s['row_begin']), int(s['row_end']))
        self._handle = h
        return
      topo = self.topo
      indptr, indices, eids, w = topo.indptr, topo.indices, topo.edge_ids, topo.edge_weights
      if self.mode == 'CUDA':
        dev = torch.device('cuda', int(self.device))
        max_id = int(indices.max().item()) if indices.numel() > 0 else 0
        idx_dtype = torch.int32 if max_id < 2 ** 31 - 1 else torch.int64
        indptr_d = indptr.to(dev)
        indices_d = indices.to(dev, dtype=idx_dtype)
        eids_d = eids.to(dev) if eids is not None else None
        w_d = w.to(dev) if w is not None else None
      else:  # ZERO_COPY: page-lock and read in place
        indptr_d, indices_d = _page_lock(indptr), _page_lock(indices)
        eids_d = _page_lock(eids) if eids is not None else None
        w_d = _page_lock(w) if w is not None else None
      h.add_shard(indptr_d, indices_d, eids_d, w_d, 0, indptr.numel() - 1)
      self._handle = h

  # ------------------------------------------------------------------ IPC
  def export_topology(self):
    return self.topo.indptr, self.topo.indices, self.topo.edge_ids, self.topo.edge_weights

  def share_ipc(self):
    with self._lock:
      if self.topo is not None:
        self.topo.share_memory_()
      return self.topo, self.mode, self.device

  @classmethod
  def from_ipc_handle(cls, ipc_handle):
    topo, mode, device = ipc_handle
    return cls(topo, mode, device)

  # ------------------------------------------------------------------ props
  @property
  def row_count(self):
    if self.topo is not None:
      return self.topo.row_count
    self.lazy_init()
    return self._handle.num_rows

  @property
  def col_count(self):
    if self._col_count is None:
      if self.topo is not None and self.topo.indices.numel() > 0:
        self._col_count = int(self.topo.indices.max().item()) + 1
      else:
        self._col_count = self.row_count
    return self._col_count

  @property
  def edge_count(self):
    return self.topo.edge_count if self.topo is not None else -1

  @property
  def graph_handler(self):
    """Native handle (GraphHandle for device modes, the Topology itself for CPU)."""
    if self.mode == 'CPU':
      return self.topo
    self.lazy_init()
    return self._handle


def rebuild_graph(ipc_handle):
  return Graph.from_ipc_handle(ipc_handle)


def reduce_graph(graph: Graph):
  return (rebuild_graph, (graph.share_ipc(),))


ForkingPickler.register(Graph, reduce_graph)
