from .partition_book import PartitionBook, GLTPartitionBook, RangePartitionBook, OffsetId2Index
from .base import (PartitionerBase, save_meta, save_node_pb, save_edge_pb, save_graph_partition,
                   save_graph_cache, save_feature_partition, save_feature_partition_chunk,
                   save_feature_partition_cache, load_partition, load_graph_partition_data,
                   load_feature_partition_data, cat_feature_cache, build_partition_feature)
from .random_partitioner import RandomPartitioner, RangePartitioner
from .frequency_partitioner import FrequencyPartitioner
# typing records and file helpers the reference exposes from `partition` (python/partition/base.py star imports)
from ..typing import FeaturePartitionData, GraphPartitionData, as_str
from ..utils import append_tensor_to_file, convert_to_tensor, ensure_dir, id2idx, load_and_concatenate_tensors
