from .partition_book import PartitionBook, GLTPartitionBook, RangePartitionBook, OffsetId2Index
from .base import (PartitionerBase, save_meta, save_node_pb, save_edge_pb, save_graph_partition,
                   save_graph_cache, save_feature_partition, save_feature_partition_chunk,
                   save_feature_partition_cache, load_partition, load_graph_partition_data,
                   load_feature_partition_data, cat_feature_cache, build_partition_feature)
from .random_partitioner import RandomPartitioner, RangePartitioner
from .frequency_partitioner import FrequencyPartitioner
