"""Property-graph *fragment* bridge (GraphScope / vineyard ingestion).

API parity: reference python/data/vineyard_utils.py:29-140 and the native reader in
v6d/vineyard_utils.cc (ToCSR, LoadVertexFeatures, LoadEdgeFeatures, GetFragVertexOffset,
GetFragVertexNum, VineyardFragHandle).  A vineyard ``ArrowFragment`` is one partition of a labelled
property graph kept as Arrow tables: a vertex table per vertex label (row i = the i-th *inner*
vertex of the fragment), an edge table per edge label, and a global-id scheme that tells which
fragment owns a vertex.  The functions below expose exactly that contract

    vineyard_to_csr(sock, fid, v_label, e_label, edge_dir, haseid)  -> (indptr, indices, eids)
    load_vertex_feature_from_vineyard(sock, fid, vcols, v_label)     -> [num_inner, F]
    load_edge_feature_from_vineyard(sock, fid, ecols, e_label)       -> [num_edges, F]
    get_frag_vertex_offset / get_frag_vertex_num / get_fid_from_gid
    VineyardPartitionBook, VineyardGid2Lid, v6d_id_select, v6d_id_filter

over pluggable *fragment backends*:

* ``ArrowDirBackend`` (built in, selected when ``sock`` is a directory): the same fragments
  materialised as Arrow IPC files -- what a GraphScope job exports, and what
  :func:`write_arrow_fragments` produces from plain tensors.  Tables are memory-mapped with
  pyarrow and converted to torch tensors column-wise without a Python loop over rows; CSR
  construction runs in the native COO->CSR op.
* a live vineyard daemon: register a backend for its socket with :func:`register_fragment_backend`
  (the ``vineyard`` Python client is not part of this image, so no client code is bundled).

Global ids are range encoded: fragment f owns the contiguous gid range of every vertex label
recorded in the store's ``meta.json`` (vineyard packs (fid, label, offset) into bit fields; only
the *mapping* matters to the loaders, and ranges keep ``gid -> fid`` a searchsorted).
"""
import json
import os
from collections.abc import Sequence
from typing import Callable, Dict, List, Optional, Tuple

import torch

from ..partition.partition_book import PartitionBook

_BACKENDS: Dict[str, Callable] = {}
_OPEN: Dict[Tuple[str, str], 'FragmentBackend'] = {}


class FragmentBackend(object):
  """What a fragment provider has to implement (all ids int64 torch tensors on the CPU)."""

  fid: int
  fnum: int

  def vertex_offset(self, v_label: str) -> int:
    raise NotImplementedError

  def vertex_num(self, v_label: str) -> int:
    raise NotImplementedError

  def vertex_ranges(self, v_label: str) -> torch.Tensor:
    """[fnum + 1] gid boundaries of the label: fragment f owns [r[f], r[f+1])."""
    raise NotImplementedError

  def vertex_columns(self, v_label: str, cols: List[str]) -> torch.Tensor:
    raise NotImplementedError

  def edge_endpoints(self, e_label: str) -> Tuple[str, str, torch.Tensor, torch.Tensor]:
    """(src vertex label, dst vertex label, src gids, dst gids) of the fragment's edges."""
    raise NotImplementedError

  def edge_columns(self, e_label: str, cols: List[str]) -> torch.Tensor:
    raise NotImplementedError


def register_fragment_backend(prefix: str, factory: Callable[[str, str], FragmentBackend]):
  """Route sockets/URIs starting with `prefix` to `factory(sock, object_id)`."""
  _BACKENDS[prefix] = factory


def _open(sock: str, object_id) -> FragmentBackend:
  key = (str(sock), str(object_id))
  if key not in _OPEN:
    backend = None
    for prefix, factory in _BACKENDS.items():
      if str(sock).startswith(prefix):
        backend = factory(str(sock), str(object_id))
        break
    if backend is None:
      path = str(sock)[len('file://'):] if str(sock).startswith('file://') else str(sock)
      if os.path.isdir(path):
        backend = ArrowDirBackend(path, str(object_id))
      else:
        raise RuntimeError(
          f'no fragment backend for {sock!r}: pass the directory of an Arrow fragment store '
          '(see write_arrow_fragments) or register a backend for a live vineyard socket with '
          'register_fragment_backend(prefix, factory)')
    _OPEN[key] = backend
  return _OPEN[key]


def _column_to_tensor(col) -> torch.Tensor:
  """One Arrow column -> [rows] or [rows, width] tensor (fixed-size lists are feature vectors)."""
  import pyarrow as pa
  arr = col.combine_chunks() if hasattr(col, 'combine_chunks') else col
  if pa.types.is_fixed_size_list(arr.type):
    width = arr.type.list_size
    flat = arr.flatten().to_numpy(zero_copy_only=False)
    return torch.from_numpy(flat.copy()).view(-1, width)
  if pa.types.is_list(arr.type) or pa.types.is_large_list(arr.type):
    flat = arr.flatten().to_numpy(zero_copy_only=False)
    rows = len(arr)
    return torch.from_numpy(flat.copy()).view(rows, -1)
  return torch.from_numpy(arr.to_numpy(zero_copy_only=False).copy())


def _columns(table, cols: List[str]) -> torch.Tensor:
  parts = []
  for c in cols:
    t = _column_to_tensor(table.column(c))
    parts.append(t.unsqueeze(1) if t.dim() == 1 else t)
  if not parts:
    return torch.empty(table.num_rows, 0)
  dtype = parts[0].dtype
  if any(p.dtype != dtype for p in parts):
    dtype = torch.float32
  return torch.cat([p.to(dtype) for p in parts], dim=1)


class ArrowDirBackend(FragmentBackend):
  """<root>/<object_id>/{meta.json, v_<label>.arrow, e_<label>.arrow}"""

  def __init__(self, root: str, object_id: str):
    import pyarrow  # noqa: F401  (fail early with a clear message)
    self.dir = os.path.join(root, object_id)
    with open(os.path.join(self.dir, 'meta.json')) as f:
      self.meta = json.load(f)
    self.fid = int(self.meta['fid'])
    self.fnum = int(self.meta['fnum'])
    self._tables = {}

  def _table(self, name: str):
    if name not in self._tables:
      import pyarrow as pa
      import pyarrow.ipc as ipc
      with pa.memory_map(os.path.join(self.dir, name + '.arrow'), 'r') as src:
        self._tables[name] = ipc.open_file(src).read_all()
    return self._tables[name]

  def vertex_ranges(self, v_label):
    return torch.tensor(self.meta['vertex_labels'][v_label]['ranges'], dtype=torch.int64)

  def vertex_offset(self, v_label):
    return int(self.meta['vertex_labels'][v_label]['ranges'][self.fid])

  def vertex_num(self, v_label):
    r = self.meta['vertex_labels'][v_label]['ranges']
    return int(r[self.fid + 1] - r[self.fid])

  def vertex_columns(self, v_label, cols):
    return _columns(self._table('v_' + v_label), list(cols))

  def edge_endpoints(self, e_label):
    m = self.meta['edge_labels'][e_label]
    t = self._table('e_' + e_label)
    return (m['src_label'], m['dst_label'], _column_to_tensor(t.column('__src__')).long(),
            _column_to_tensor(t.column('__dst__')).long())

  def edge_columns(self, e_label, cols):
    return _columns(self._table('e_' + e_label), list(cols))


# ----------------------------------------------------------------------------- reference API
def vineyard_to_csr(sock, fid, v_label_name, e_label_name, edge_dir, haseid=0):
  """CSR of one edge label over the fragment's inner vertices of `v_label_name`.

  Rows are LOCAL vertex offsets (gid - fragment offset), column entries are GLOBAL ids;
  `edge_dir='out'` keys rows by source, `'in'` by destination (i.e. the result is then the CSC
  of the fragment).  Returns (indptr, indices, edge_ids | None); edge ids are the row numbers of
  the fragment's edge table, which is what load_edge_feature_from_vineyard is indexed by.
  """
  from ..utils.topo import coo_to_csr  # native COO -> CSR (csrc/cpu/cpu_ops.cc)
  fr = _open(sock, fid)
  src_label, dst_label, src, dst = fr.edge_endpoints(e_label_name)
  key_label = src_label if edge_dir == 'out' else dst_label
  assert key_label == v_label_name, f'edge label {e_label_name} is keyed by {key_label}, not {v_label_name}'
  offset, num = fr.vertex_offset(v_label_name), fr.vertex_num(v_label_name)
  rows = (src if edge_dir == 'out' else dst) - offset
  cols = dst if edge_dir == 'out' else src
  inner = (rows >= 0) & (rows < num)
  eids = torch.arange(rows.numel(), dtype=torch.int64)
  if not bool(inner.all()):
    rows, cols, eids = rows[inner], cols[inner], eids[inner]
  indptr, indices, out_eids, _ = coo_to_csr(rows, cols, eids, None, node_sizes=(num, num))
  return indptr, indices, (out_eids if haseid else None)


def load_vertex_feature_from_vineyard(sock, fid, vcols, v_label_name):
  return _open(sock, fid).vertex_columns(v_label_name, vcols)


def load_edge_feature_from_vineyard(sock, fid, ecols, e_label_name):
  return _open(sock, fid).edge_columns(e_label_name, ecols)


def get_frag_vertex_offset(sock, fid, v_label_name):
  return _open(sock, fid).vertex_offset(v_label_name)


def get_frag_vertex_num(sock, fid, v_label_name):
  return _open(sock, fid).vertex_num(v_label_name)


def get_fid_from_gid(gid, sock=None, fid=None, v_label_name=None):
  """Fragment id that owns `gid` (needs the store to know the ranges)."""
  assert sock is not None and fid is not None and v_label_name is not None, \
    'range-encoded gids need the fragment store: pass sock, fid and v_label_name'
  ranges = _open(sock, fid).vertex_ranges(v_label_name)
  g = torch.as_tensor(gid, dtype=torch.int64)
  out = torch.searchsorted(ranges, g.reshape(-1), right=True) - 1
  return out.reshape(g.shape) if g.dim() else int(out.item())


class VineyardPartitionBook(PartitionBook):
  """gid -> partition through the fragment ranges (optionally remapped by `fid2pid`)."""

  def __init__(self, sock, obj_id, v_label_name, fid2pid: Optional[Dict[int, int]] = None):
    self._sock, self._obj_id, self._v_label_name = sock, obj_id, v_label_name
    fr = _open(sock, obj_id)
    self._ranges = fr.vertex_ranges(v_label_name)
    self._offset = fr.vertex_offset(v_label_name)
    self._fid2pid = fid2pid

  def __getitem__(self, gids) -> torch.Tensor:
    fids = self.gid2fid(gids)
    if self._fid2pid is not None:
      lut = torch.full((int(self._ranges.numel()) - 1,), -1, dtype=torch.int64)
      for f, p in self._fid2pid.items():
        lut[int(f)] = int(p)
      return lut[fids].to(torch.int32)
    return fids.to(torch.int32)

  @property
  def device(self):
    return torch.device('cpu')

  @property
  def offset(self):
    return self._offset

  def gid2fid(self, gids) -> torch.Tensor:
    g = torch.as_tensor(gids, dtype=torch.int64).cpu()
    return torch.searchsorted(self._ranges, g, right=True) - 1

  def id_filter(self, node_pb, partition_idx):
    return v6d_id_filter(self, partition_idx)


class VineyardGid2Lid(Sequence):
  """id2index for Feature: gid -> row of the fragment's vertex table."""

  def __init__(self, sock, fid, v_label_name):
    self._offset = get_frag_vertex_offset(sock, fid, v_label_name)
    self._vnum = get_frag_vertex_num(sock, fid, v_label_name)

  def __getitem__(self, gids):
    return gids - self._offset

  def __len__(self):
    return self._vnum


def v6d_id_select(srcs, p_mask, node_pb: PartitionBook):
  """Inner vertices among `srcs` selected by `p_mask`, as local offsets of the partition."""
  return torch.masked_select(srcs, p_mask) - node_pb.offset


def v6d_id_filter(node_pb: VineyardPartitionBook, partition_idx):
  """Global ids of the inner vertices of this fragment."""
  fr = _open(node_pb._sock, node_pb._obj_id)
  off, num = fr.vertex_offset(node_pb._v_label_name), fr.vertex_num(node_pb._v_label_name)
  return torch.arange(off, off + num, dtype=torch.int64)


# ----------------------------------------------------------------------------- writer
def _to_arrow_column(t: torch.Tensor):
  import pyarrow as pa
  a = t.detach().cpu().contiguous().numpy()
  if a.ndim == 1:
    return pa.array(a)
  return pa.FixedSizeListArray.from_arrays(pa.array(a.reshape(-1)), a.shape[1])


def write_arrow_fragments(root: str, name: str, num_fragments: int,
                          vertices: Dict[str, Dict[str, torch.Tensor]],
                          edges: Dict[str, Tuple[str, str, torch.Tensor, Dict[str, torch.Tensor]]],
                          edge_owner: str = 'src') -> List[str]:
  """Range-partition a labelled property graph into `num_fragments` Arrow fragments.

  vertices: {v_label: {column: tensor[num_vertices(, width)]}}   (vertex i has gid i)
  edges:    {e_label: (src_label, dst_label, edge_index[2, E], {column: tensor[E(, width)]})}
  Every fragment gets the edges whose `edge_owner` endpoint it owns ('src' -> usable with
  edge_dir='out', 'dst' -> edge_dir='in').  Returns the object ids ("<name>_<fid>").
  """
  import pyarrow as pa
  import pyarrow.ipc as ipc
  ranges = {}
  for vl, cols in vertices.items():
    n = next(iter(cols.values())).shape[0]
    per = (n + num_fragments - 1) // num_fragments
    ranges[vl] = [min(f * per, n) for f in range(num_fragments + 1)]
  ids = []
  for f in range(num_fragments):
    oid = f'{name}_{f}'
    d = os.path.join(root, oid)
    os.makedirs(d, exist_ok=True)
    meta = {'fid': f, 'fnum': num_fragments, 'vertex_labels': {}, 'edge_labels': {}}
    for vl, cols in vertices.items():
      b, e = ranges[vl][f], ranges[vl][f + 1]
      table = pa.table({c: _to_arrow_column(t[b:e]) for c, t in cols.items()})
      with ipc.new_file(os.path.join(d, f'v_{vl}.arrow'), table.schema) as w:
        w.write_table(table)
      meta['vertex_labels'][vl] = {'ranges': ranges[vl]}
    for el, (sl, dl, ei, cols) in edges.items():
      own_label = sl if edge_owner == 'src' else dl
      own = ei[0] if edge_owner == 'src' else ei[1]
      b, e = ranges[own_label][f], ranges[own_label][f + 1]
      sel = ((own >= b) & (own < e)).nonzero(as_tuple=True)[0]
      data = {'__src__': _to_arrow_column(ei[0][sel]), '__dst__': _to_arrow_column(ei[1][sel])}
      for c, t in cols.items():
        data[c] = _to_arrow_column(t[sel])
      table = pa.table(data)
      with ipc.new_file(os.path.join(d, f'e_{el}.arrow'), table.schema) as w:
        w.write_table(table)
      meta['edge_labels'][el] = {'src_label': sl, 'dst_label': dl, 'num_edges': int(sel.numel())}
    with open(os.path.join(d, 'meta.json'), 'w') as fjson:
      json.dump(meta, fjson)
    ids.append(oid)
  return ids
