from .data import Data, HeteroData
from .transform import to_data, to_hetero_data
from .node_loader import NodeLoader, SeedBatcher
from .neighbor_loader import NeighborLoader
from .link_loader import LinkLoader, get_edge_label_index
from .link_neighbor_loader import LinkNeighborLoader
from .subgraph_loader import SubGraphLoader
# sampler-side types are importable from `loader` as in the reference (python/loader/__init__.py star imports)
from ..data.dataset import Dataset
from ..sampler import (BaseSampler, EdgeSamplerInput, HeteroSamplerOutput, NegativeSampling, NeighborSampler,
                       NodeSamplerInput, SamplerOutput)
from ..typing import reverse_edge_type
from ..utils import convert_to_tensor
