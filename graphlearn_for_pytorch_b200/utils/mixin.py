"""`CastMixin.cast`: coerce "something that describes an instance" into an instance.

Sampler inputs / outputs accept the instance itself, a positional tuple or list, a keyword mapping, or plain
constructor arguments, so that API entry points can take whatever the caller has at hand
(capability parity: reference utils/mixin.py).
"""
from collections.abc import Mapping, Sequence


class CastMixin(object):
  @classmethod
  def cast(cls, *values, **fields):
    """cast(obj) | cast((a, b, ...)) | cast({'a': ..}) | cast(a, b, x=..) -> instance of `cls` (None stays None)."""
    single = values[0] if (len(values) == 1 and not fields) else _NOTHING
    if single is _NOTHING:
      return cls(*values, **fields)
    if single is None or isinstance(single, CastMixin):
      return single
    if isinstance(single, Mapping):
      return cls(**single)
    if isinstance(single, Sequence) and not isinstance(single, (str, bytes)):
      return cls(*single)
    return cls(single)


_NOTHING = object()
