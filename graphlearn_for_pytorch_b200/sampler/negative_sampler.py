"""RandomNegativeSampler (API parity: reference python/sampler/negative_sampler.py:21-57).

CPU: native C++ (csrc/cpu/cpu_ops.cc); GPU: fused draw/reject/compact kernel
(csrc/cuda/graph_ops.cu).  Row and column draws use independent Philox words.
"""
import threading

import torch

from ..ops import require_native
from ..utils.common import seed_everything  # noqa: F401

_counter_lock = threading.Lock()
_counter = [0]


def _next_stream() -> int:
  with _counter_lock:
    _counter[0] = (_counter[0] + 1) & 0x3FFFFFFF
    return _counter[0]


class RandomNegativeSampler(object):
  """Sample (row, col) pairs that are *not* edges of `graph`.

  Args:
    graph: data.Graph.
    mode: 'CUDA' or 'CPU'.
    edge_dir: 'out' (CSR) or 'in' (CSC: rows/cols are swapped on output).
  """

  def __init__(self, graph, mode: str = 'CUDA', edge_dir: str = 'out', seed=None):
    self.graph = graph
    self.mode = 'CPU' if graph.mode == 'CPU' else str(mode).upper()
    self.edge_dir = edge_dir
    self.seed = seed if seed is not None else int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    self._nat = require_native()
    if self.mode != 'CPU':
      graph.lazy_init()

  def sample(self, req_num: int, trials_num: int = 5, padding: bool = False) -> torch.Tensor:
    """Returns a [2, n] tensor, n <= req_num (== req_num when padding=True)."""
    stream = _next_stream()
    n_rows, n_cols = self.graph.row_count, self.graph.col_count
    if self.mode == 'CPU':
      topo = self.graph.topo
      rows, cols = self._nat.cpu_negative_sample(topo.indptr, topo.indices, n_rows, n_cols, int(req_num),
                                                 int(trials_num), bool(padding), True, self.seed, stream)
    else:
      h = self.graph.graph_handler
      rows, cols, count = h.negative_sample(n_rows, n_cols, int(req_num), int(trials_num), bool(padding),
                                            self.seed, stream)
      n = int(req_num) if padding else int(count.item())
      rows, cols = rows[:n], cols[:n]
    if self.edge_dir == 'in':
      rows, cols = cols, rows
    return torch.stack([rows, cols])
