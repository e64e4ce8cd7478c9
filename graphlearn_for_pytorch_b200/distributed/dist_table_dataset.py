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
      off += e.shape[1]
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
    if isinstance(nf, dict) and not nf:
      nf = nids = None
    return cls(output_dir, num_nodes, ei, eids, nf, nids, **kwargs)


class DistTableDataset(DistDataset):
  """Every worker reads a SLICE of the node / edge tables; the workers partition the graph online together
  (`DistTableRandomPartitioner` over RPC) and each loads its partition (reference:
  python/distributed/dist_table_dataset.py:30-250)."""
  def load(self, num_nodes=0, edge_tables=None, node_tables=None, graph_mode: str = 'CPU', feature_with_gpu=False,
           label_col: Optional[str] = 'label', id_col: str = 'id', device=None, *, num_partitions: Optional[int] = None,
           partition_idx: Optional[int] = None, device_group_list=None, reader_threads: int = 10,
           reader_capacity: int = 10240, reader_batch_size: int = 1024, label: Optional[str] = None,
           edge_assign_strategy: str = 'by_src', chunk_size: int = 10000,
           master_addr: Optional[str] = None, master_port: Optional[int] = None, num_rpc_threads: int = 16,
           **kwargs):
    """Collectively partition the table slices and load this rank's partition.

    Keyword names follow the reference (`dist_table_dataset.py:36-56`): with `master_addr` / `master_port` (and
    `num_partitions`, `partition_idx`) the worker group and the RPC agent are brought up here if they are not yet;
    `label` is an alias of `label_col`; `edge_assign_strategy` / `chunk_size` go to the partitioner;
    the `reader_*` knobs of the ODPS reader are accepted and unused (tables are read through pyarrow in one pass)."""
    from .rpc import all_gather, init_rpc, rpc_is_initialized
    del reader_threads, reader_capacity, reader_batch_size, device_group_list
    kwargs.setdefault('edge_assign_strategy', edge_assign_strategy)
    kwargs.setdefault('chunk_size', chunk_size)
    if label is not None:
      label_col = label
    if get_context() is None and num_partitions is not None:
      from .dist_context import init_worker_group
      init_worker_group(int(num_partitions), int(partition_idx or 0), 'table-dataset')
    if not rpc_is_initialized() and master_addr is not None:
      init_rpc(master_addr, int(master_port), num_rpc_threads=num_rpc_threads)
    ctx = get_context()
    assert ctx is not None and rpc_is_initialized(), 'init_worker_group() + init_rpc() first (or pass master_addr)'
    out = kwargs.pop('output_dir', None)
    if out is None:
      # every rank must write under the same root: rank 0 picks it
      mine = tempfile.mkdtemp(prefix='glt_b200_table_') if ctx.rank == 0 else None
      out = [v for v in all_gather(mine).values() if v is not None][0]
    # dense global edge ids: this rank's slice starts after the slices of the lower ranks
    n_local = sum(len(_read_table(src)[kwargs.get('src_col', 'src_id')]) for src in edge_tables.values())
    counts = all_gather((ctx.rank, n_local))
    offset = sum(n for r, n in counts.values() if r < ctx.rank)
    part = DistTableRandomPartitioner.from_tables(out, num_nodes, edge_tables, node_tables, id_col=id_col,
                                                  edge_id_offset=offset, **kwargs)
    part.partition()
    DistDataset.load(self, out, ctx.rank, graph_mode=graph_mode, feature_with_gpu=feature_with_gpu, device=device)
    if label_col:
      # labels travel with the node tables: gather every rank's (id, label) slice
      import numpy as np
      mine = {}
      for nt, src in (node_tables or {}).items():
        cols = _read_table(src)
        if label_col in cols:
          mine[nt] = (torch.from_numpy(cols[id_col].astype(np.int64)), torch.from_numpy(cols[label_col].astype(np.int64)))
      gathered = all_gather(mine)
      labels = {}
      for part_labels in gathered.values():
        for nt, (ids, lab) in part_labels.items():
          n = num_nodes[nt] if isinstance(num_nodes, dict) else num_nodes
          full = labels.setdefault(nt, torch.full((n,), -1, dtype=torch.int64))
          full[ids] = lab
      if labels:
        self.init_node_labels(labels if isinstance(num_nodes, dict) else next(iter(labels.values())))
    return self
