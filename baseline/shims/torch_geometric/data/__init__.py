import torch


class Data(object):
  def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kwargs):
    self.x, self.edge_index, self.edge_attr, self.y = x, edge_index, edge_attr, y
    for k, v in kwargs.items():
      setattr(self, k, v)

  def __getitem__(self, key):
    return getattr(self, key, None)

  def __setitem__(self, key, value):
    setattr(self, key, value)

  def to(self, device):
    for k, v in list(self.__dict__.items()):
      if isinstance(v, torch.Tensor):
        setattr(self, k, v.to(device))
    return self


class _Store(dict):
  __getattr__ = dict.get
  __setattr__ = dict.__setitem__


class HeteroData(object):
  def __init__(self, **kwargs):
    self._stores = {}
    for k, v in kwargs.items():
      setattr(self, k, v)

  def __getitem__(self, key):
    return self._stores.setdefault(key, _Store())

  def __setitem__(self, key, value):
    self._stores[key] = value
