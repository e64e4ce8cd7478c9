from .graph import Topology, Graph
from .unified_tensor import UnifiedTensor
from .feature import DeviceGroup, Feature
from .reorder import sort_by_in_degree
from .dataset import Dataset, random_split
from .table_dataset import TableDataset
