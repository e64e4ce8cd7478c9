"""Minimal PyG-compatible `Data` / `HeteroData` containers.

The reference returns torch_geometric objects (python/loader/transform.py:20); this
framework must not depend on PyG, so it ships attribute-compatible containers
(x, y, edge_index, edge_attr, node, edge, batch, batch_size, num_sampled_nodes, ...;
`data['paper'].x`, `data['a','r','b'].edge_index`, `x_dict`, `edge_index_dict`).
If torch_geometric is importable and GLT_B200_USE_PYG=1, real PyG objects are used.
"""
import os
from typing import Any, Dict

import torch

_USE_PYG = False
if os.environ.get('GLT_B200_USE_PYG', '0') == '1':
  try:  # pragma: no cover - PyG is not installed in the build image
    from torch_geometric.data import Data as _PygData, HeteroData as _PygHeteroData
    _USE_PYG = True
  except ImportError:
    _USE_PYG = False


class _Storage(object):
  """Attribute bag with dict access and tensor-wise `.to()`."""

  def __init__(self, **kwargs):
    object.__setattr__(self, '_store', {})
    for k, v in kwargs.items():
      self._store[k] = v

  def __getattr__(self, key):
    store = object.__getattribute__(self, '_store')
    if key in store:
      return store[key]
    if key.startswith('__'):
      raise AttributeError(key)
    return None

  def __setattr__(self, key, value):
    self._store[key] = value

  def __getitem__(self, key):
    return self._store.get(key)

  def __setitem__(self, key, value):
    self._store[key] = value

  def __contains__(self, key):
    return key in self._store and self._store[key] is not None

  def keys(self):
    return [k for k, v in self._store.items() if v is not None]

  def items(self):
    return [(k, v) for k, v in self._store.items() if v is not None]

  def to_dict(self) -> Dict[str, Any]:
    return dict(self.items())

  def _apply(self, fn):
    def rec(v):
      if isinstance(v, torch.Tensor):
        return fn(v)
      if isinstance(v, dict):
        return {k: rec(x) for k, x in v.items()}
      if isinstance(v, (list, tuple)):
        return type(v)(rec(x) for x in v)
      return v
    for k in list(self._store.keys()):
      self._store[k] = rec(self._store[k])
    return self

  def to(self, device, non_blocking: bool = False):
    return self._apply(lambda t: t.to(device, non_blocking=non_blocking))

  def cpu(self):
    return self.to('cpu')

  def cuda(self, device=None):
    return self.to(device if device is not None else 'cuda')

  def pin_memory(self):
    return self._apply(lambda t: t.pin_memory() if t.device.type == 'cpu' else t)

  def __repr__(self):
    def fmt(v):
      if isinstance(v, torch.Tensor):
        return list(v.shape)
      return v if not isinstance(v, dict) else '{...}'
    body = ', '.join(f'{k}={fmt(v)}' for k, v in self.items())
    return f'{self.__class__.__name__}({body})'


class Data(_Storage):
  """Homogeneous mini-batch."""

  def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kwargs):
    super().__init__(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, **kwargs)

  @property
  def num_nodes(self):
    if self.node is not None:
      return int(self.node.numel())
    if self.x is not None:
      return int(self.x.shape[0])
    return int(self.edge_index.max()) + 1 if self.edge_index is not None and self.edge_index.numel() else 0

  @property
  def num_edges(self):
    return int(self.edge_index.shape[1]) if self.edge_index is not None else 0


class HeteroData(_Storage):
  """Heterogeneous mini-batch: per-node-type and per-edge-type storages."""

  def __init__(self, **kwargs):
    super().__init__(**kwargs)
    object.__setattr__(self, '_node_stores', {})
    object.__setattr__(self, '_edge_stores', {})

  def __getitem__(self, key):
    if isinstance(key, tuple) and len(key) == 3:
      return self._edge_stores.setdefault(tuple(key), _Storage())
    if isinstance(key, str) and key in self._store:
      return self._store[key]
    if isinstance(key, str):
      return self._node_stores.setdefault(key, _Storage())
    raise KeyError(key)

  def __setitem__(self, key, value):
    if isinstance(key, str):
      self._store[key] = value
    else:
      raise KeyError(key)

  @property
  def node_types(self):
    return list(self._node_stores.keys())

  @property
  def edge_types(self):
    return list(self._edge_stores.keys())

  def metadata(self):
    return self.node_types, self.edge_types

  def _collect(self, stores, attr):
    return {k: s[attr] for k, s in stores.items() if s[attr] is not None}

  @property
  def x_dict(self):
    return self._collect(self._node_stores, 'x')

  @property
  def y_dict(self):
    return self._collect(self._node_stores, 'y')

  @property
  def node_dict(self):
    return self._collect(self._node_stores, 'node')

  @property
  def edge_index_dict(self):
    return self._collect(self._edge_stores, 'edge_index')

  @property
  def edge_attr_dict(self):
    return self._collect(self._edge_stores, 'edge_attr')

  def to(self, device, non_blocking: bool = False):
    super().to(device, non_blocking)
    for s in list(self._node_stores.values()) + list(self._edge_stores.values()):
      s.to(device, non_blocking)
    return self

  def __repr__(self):
    parts = [f'{k}={v!r}' for k, v in self._node_stores.items()]
    parts += [f'{k}={v!r}' for k, v in self._edge_stores.items()]
    return 'HeteroData(' + ', '.join(parts) + ')'


if _USE_PYG:  # pragma: no cover
  Data, HeteroData = _PygData, _PygHeteroData
