"""CastMixin: build a dataclass from itself / a tuple / a dict (reference utils/mixin.py)."""


class CastMixin:
  @classmethod
  def cast(cls, *args, **kwargs):
    if len(args) == 1 and len(kwargs) == 0:
      elem = args[0]
      if elem is None:
        return None
      if isinstance(elem, CastMixin):
        return elem
      if isinstance(elem, (tuple, list)):
        return cls(*elem)
      if isinstance(elem, dict):
        return cls(**elem)
    return cls(*args, **kwargs)
