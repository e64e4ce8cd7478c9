from .base import *
from .negative_sampler import RandomNegativeSampler
from .neighbor_sampler import NeighborSampler
# helper names the reference's `sampler` namespace carries along (python/sampler/__init__.py star imports)
from ..data.graph import Graph
from ..typing import reverse_edge_type
from ..utils import count_dict, format_hetero_sampler_output, id2idx, merge_dict, merge_hetero_sampler_output


def __getattr__(name):
  # `pywrap`: the native-module handle of the reference's Python layer (`from .. import py_graphlearn_torch as
  # pywrap`); resolved on first use because the facade imports this sub-package
  if name == 'pywrap':
    import importlib
    return importlib.import_module('..py_graphlearn_torch', __name__)
  raise AttributeError(name)
