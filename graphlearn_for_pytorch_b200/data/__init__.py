from .graph import Topology, Graph
from .unified_tensor import UnifiedTensor
from .feature import DeviceGroup, Feature
from .reorder import sort_by_in_degree
from .dataset import Dataset, random_split
from .table_dataset import TableDataset
from . import vineyard_utils
from .vineyard_utils import (vineyard_to_csr, load_vertex_feature_from_vineyard, load_edge_feature_from_vineyard,
                             get_fid_from_gid, get_frag_vertex_offset, get_frag_vertex_num, VineyardPartitionBook,
                             VineyardGid2Lid, v6d_id_select, v6d_id_filter, write_arrow_fragments,
                             register_fragment_backend)
from .quantize import (quantize_mxfp8, dequantize_mxfp8, mxfp8_row_bytes, quantize_mxfp8_parts,
                       pack_mx_scale_blocks)
