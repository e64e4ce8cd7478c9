"""graphlearn_for_pytorch_b200 -- a Blackwell-native GNN sampling / feature engine with the
capabilities and PyG-compatible API surface of alibaba/graphlearn-for-pytorch.

Layer map (bottom-up): csrc (C++ / sm_100a CUDA) -> ops -> data -> sampler -> loader ->
channel / partition -> parallel (NVLink symmetric heap, P2P kernels) -> distributed ->
models (GraphSAGE / R-GCN engines).
"""
import os as _os

_os.environ.setdefault('TORCH_CPP_LOG_LEVEL', 'ERROR')

from . import typing  # noqa: E402
from . import utils  # noqa: E402
from . import ops  # noqa: E402
from . import data  # noqa: E402
from . import sampler  # noqa: E402
from . import loader  # noqa: E402
from .typing import *  # noqa: E402,F401,F403  (NodeType, EdgeType, InputNodes, Split, ... at top level, as in the reference)

__version__ = '0.1.0'


def _lazy(name):
  import importlib
  return importlib.import_module(f'{__name__}.{name}')


def __getattr__(name):
  if name in ('channel', 'partition', 'distributed', 'parallel', 'models', 'py_graphlearn_torch'):
    return _lazy(name)
  raise AttributeError(name)
