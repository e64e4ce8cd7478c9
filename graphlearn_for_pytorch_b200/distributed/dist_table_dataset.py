"""DistTableDataset / DistTableRandomPartitioner: build an in-memory distributed dataset from
table slices (parity: reference python/distributed/dist_table_dataset.py:40-360, which reads
Alibaba ODPS tables through `common_io`; here each rank reads its slice from parquet / csv /
in-memory columns via data.table_dataset._read_table)."""
import tempfile
from typing import Dict, Optional

import numpy as np
import torch

from ..data.table_dataset import _parse_feature_column, _read_table
from ..typing import EdgeType, NodeType
from .dist_context import get_context
from .dist_dataset import DistDataset
from .dist_random_partitioner import DistRandomPartitioner


class DistTableRandomPartitioner(DistRandomPartitioner):
  """DistRandomPartitioner whose inputs come from this rank's table slices."""

  @classmethod
  def from_tables(cls, output_dir: str, num_nodes, edge_tables: Dict[Optional[EdgeType], object],
                  node_tables: Dict[Optional[NodeType], object], src_col='src_id', dst_col='dst_id',
                  id_col='id', feature_col='feature', edge_id_offset: int = 0, **kwargs):
    hetero = isinstance(num_nodes, dict)
    ei, eids, nf, nids = {}, {}, {}, {}
    off = edge_id_offset
    for et, src in edge_tables.items():
      cols = _read_table(src)
      e = torch.stack([torch.from_numpy(cols[src_col].astype(np.int64)),
                       torch.from_numpy(cols[dst_col].astype(np.int64))])
      ei[et] = e
      eids[et] = torch.arange(off, off + e.shape[1]) if 'edge_id' not in cols else \
          torch.from_numpy(cols['edge_id'].astype(np.int64))
    for nt, src in (node_tables or {}).items():
      cols = _read_table(src)
      nids[nt] = torch.from_numpy(cols[id_col].astype(np.int64))
      nf[nt] = _parse_feature_column(cols[feature_col])
    if not hetero:
      k = next(iter(ei))
      ei, eids = ei[k], eids[k]
      if nf:
        kn = next(iter(nf))
        nf, nids = nf[kn], nids[kn]
      else:
        nf = nids = None
    return cls(output_dir, num_nodes, ei, eids, nf or None, nids or None, **kwargs)


class DistTableDataset(DistDataset):
  def load(self, num_nodes, edge_tables, node_tables=None, graph_mode: str = 'CPU', feature_with_gpu=False,
           label_col: Optional[str] = 'label', id_col: str = 'id', device=None, **kwargs):
    """Collectively partition the table slices and load this rank's partition."""
    ctx = get_context()
    assert ctx is not None, 'init_worker_group() + init_rpc() first'
    out = kwargs.pop('output_dir', None) or tempfile.mkdtemp(prefix='glt_b200_table_')
    part = DistTableRandomPartitioner.from_tables(out, num_nodes, edge_tables, node_tables, id_col=id_col, **kwargs)
    part.partition()
    DistDataset.load(self, out, ctx.rank, graph_mode=graph_mode, feature_with_gpu=feature_with_gpu, device=device)
    return self
