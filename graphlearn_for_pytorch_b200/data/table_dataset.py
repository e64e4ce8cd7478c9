"""TableDataset: build a Dataset from tabular sources.

The reference reads Alibaba ODPS tables through `common_io`
(python/data/table_dataset.py:82-144), which does not exist outside Alibaba
Cloud.  The same surface is kept, but rows come from any iterable "table reader":
pyarrow/parquet/CSV files, or a callable yielding record batches.
"""
from typing import Callable, Dict, Iterable, Optional, Tuple, Union

import numpy as np
import torch

from ..typing import EdgeType, NodeType
from .dataset import Dataset


def _read_table(src, columns=None) -> Dict[str, np.ndarray]:
  """Read a whole table into numpy columns.  `src` may be a path (.parquet/.csv),
  a pyarrow Table, a dict of arrays or a callable returning one of those."""
  if callable(src):
    src = src()
  if isinstance(src, dict):
    return {k: np.asarray(v) for k, v in src.items()}
  try:
    import pyarrow as pa
    import pyarrow.csv as pacsv
    import pyarrow.parquet as pq
  except ImportError as e:  # pragma: no cover
    raise RuntimeError('pyarrow is required to read table files') from e
  if isinstance(src, str):
    table = pq.read_table(src, columns=columns) if src.endswith('.parquet') else pacsv.read_csv(src)
  else:
    table = src
  return {name: table.column(name).to_numpy(zero_copy_only=False) for name in table.column_names}


def _parse_feature_column(col: np.ndarray, sep=':') -> torch.Tensor:
  """'0.1:0.2:...' strings (ODPS convention) or list columns -> [N, F] float tensor."""
  if col.dtype == object:
    first = col[0]
    if isinstance(first, (bytes, str)):
      rows = [np.asarray((c.decode() if isinstance(c, bytes) else c).split(sep), dtype=np.float32) for c in col]
    else:
      rows = [np.asarray(c, dtype=np.float32) for c in col]
    return torch.from_numpy(np.stack(rows))
  return torch.from_numpy(np.asarray(col, dtype=np.float32)).reshape(len(col), -1)


class TableDataset(Dataset):
  """`Dataset` filled from tables: edge tables `(src_id, dst_id[, weight])` and node tables `(id, feature[, label])`
  given as parquet / CSV paths, pyarrow tables, dicts of columns or callables (reference: ODPS tables through
  common_io, python/data/table_dataset.py:27-144)."""
  def load(self, edge_tables: Optional[Dict[EdgeType, object]] = None,
           node_tables: Optional[Dict[NodeType, object]] = None,
           graph_mode: str = 'ZERO_COPY', sort_func: Optional[Callable] = None,
           split_ratio: float = 0.0, device_group_list=None, directed: bool = True,
           label: Optional[str] = 'label', device: Optional[int] = None,
           src_col: str = 'src_id', dst_col: str = 'dst_id', id_col: str = 'id',
           feature_col: str = 'feature', weight_col: Optional[str] = None, reader_threads: int = 10,
           reader_capacity: int = 10240, reader_batch_size: int = 1024, **kwargs):
    """edge tables need (src_id, dst_id[, weight]); node tables need (id, feature[, label]).  The `reader_*`
    arguments of the reference's ODPS reader (table_dataset.py:36-39) are accepted and unused: parquet / CSV /
    pyarrow tables are read in one pass."""
    del reader_threads, reader_capacity, reader_batch_size
    assert edge_tables, 'at least one edge table is required'
    hetero = len(edge_tables) > 1 or (node_tables is not None and len(node_tables) > 1)
    edge_index, edge_weights = {}, {}
    for et, src in edge_tables.items():
      cols = _read_table(src)
      ei = torch.stack([torch.from_numpy(cols[src_col].astype(np.int64)),
                        torch.from_numpy(cols[dst_col].astype(np.int64))])
      if not directed:
        ei = torch.cat([ei, ei.flip(0)], dim=1)
      edge_index[et] = ei
      if weight_col and weight_col in cols:
        w = torch.from_numpy(cols[weight_col].astype(np.float32))
        edge_weights[et] = w if directed else torch.cat([w, w])
    feats, labels = {}, {}
    for nt, src in (node_tables or {}).items():
      cols = _read_table(src)
      ids = torch.from_numpy(cols[id_col].astype(np.int64))
      f = _parse_feature_column(cols[feature_col])
      full = torch.zeros(int(ids.max()) + 1, f.shape[1])
      full[ids] = f
      feats[nt] = full
      if label and label in cols:
        lab = torch.full((full.shape[0],), -1, dtype=torch.int64)
        lab[ids] = torch.from_numpy(cols[label].astype(np.int64))
        labels[nt] = lab
    if hetero:
      self.init_graph(edge_index, edge_weights=edge_weights or None, graph_mode=graph_mode,
                      directed=directed, device=device)
      self.init_node_features(feats or None, sort_func=sort_func, split_ratio=split_ratio,
                              device_group_list=device_group_list, device=device)
      self.init_node_labels(labels or None)
    else:
      et = next(iter(edge_index))
      self.init_graph(edge_index[et], edge_weights=edge_weights.get(et), graph_mode=graph_mode,
                      directed=directed, device=device)
      if feats:
        nt = next(iter(feats))
        self.init_node_features(feats[nt], sort_func=sort_func, split_ratio=split_ratio,
                                device_group_list=device_group_list, device=device)
        if nt in labels:
          self.init_node_labels(labels[nt])
    return self


def rebuild_table_dataset(ipc_handle):
  ds = TableDataset.from_ipc_handle(ipc_handle)
  return ds


def reduce_table_dataset(dataset: TableDataset):
  return (rebuild_table_dataset, (dataset.share_ipc(),))


from multiprocessing.reduction import ForkingPickler  # noqa: E402
ForkingPickler.register(TableDataset, reduce_table_dataset)
