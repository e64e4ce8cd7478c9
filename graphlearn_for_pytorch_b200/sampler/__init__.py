from .base import *
from .negative_sampler import RandomNegativeSampler
from .neighbor_sampler import NeighborSampler
# helper names the reference's `sampler` namespace carries along (python/sampler/__init__.py star imports)
from ..data.graph import Graph
from ..typing import reverse_edge_type
from ..utils import count_dict, format_hetero_sampler_output, id2idx, merge_dict, merge_hetero_sampler_output
