from .peer import exchange_peer_tensors, exchange_objects, range_bounds, world_info
from .partitioned import (PartitionedGraph, PartitionedFeature, shard_topology, partition_hetero_graph,
                          hotness_balanced_order)
