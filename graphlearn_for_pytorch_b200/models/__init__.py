from .sage import SAGEConv, GraphSAGE, GraphSageEngine
from .trainer import GraphSageTrainer
from .rgnn import RGNN, RelSAGEConv, RelGCNConv, RelGATConv
from .seal import drnl_node_labeling, DGCNN
from .hgt import HGT, HGTConv
from .hetero_engine import HeteroSageEngine
