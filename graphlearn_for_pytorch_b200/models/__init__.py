from .sage import SAGEConv, GraphSAGE, GraphSageEngine
