"""Utilities: every public name of the sub-modules is re-exported here (`glt.utils.get_free_port`, ...)."""
import importlib

_SUBMODULES = ('tensor', 'common', 'device', 'units', 'exit_status', 'mixin', 'singleton', 'topo', 'synthetic', 'tracing',
               'ogb_io')

for _name in _SUBMODULES:
  _mod = importlib.import_module(f'{__name__}.{_name}')
  _public = getattr(_mod, '__all__', None) or [k for k in vars(_mod) if not k.startswith('_')]
  for _k in _public:
    globals()[_k] = getattr(_mod, _k)
del _name, _mod, _public, _k

# the reference's star-import chain also exposes this typing helper as `glt.utils.reverse_edge_type`
from ..typing import reverse_edge_type  # noqa: E402,F401
