"""Public API surface: every name / constructor argument a user of the reference relies on
must exist here with the same spelling."""
import inspect

import graphlearn_for_pytorch_b200 as glt


def _params(fn):
  return list(inspect.signature(fn).parameters)


def _has(obj, names):
  missing = [n for n in names if not hasattr(obj, n)]
  assert not missing, f'{obj}: missing {missing}'


def test_data_api():
  d = glt.data
  assert _params(d.Topology.__init__)[1:6] == ['edge_index', 'edge_ids', 'edge_weights', 'input_layout', 'layout']
  assert _params(d.Graph.__init__)[1:4] == ['topo', 'mode', 'device']
  assert _params(d.DeviceGroup.__init__)[1:3] == ['group_id', 'device_list']
  assert _params(d.Feature.__init__)[1:8] == ['feature_tensor', 'id2index', 'split_ratio', 'device_group_list', 'device',
                                              'with_gpu', 'dtype']
  _has(d.Feature, ['__getitem__', 'cpu_get', 'shape', 'size', 'share_ipc', 'from_ipc_handle'])
  assert _params(d.UnifiedTensor.__init__)[1:3] == ['current_device', 'dtype']
  _has(d.UnifiedTensor, ['init_from', 'append_shared_tensor', 'append_cpu_tensor', 'share_ipc', 'new_from_ipc'])
  assert _params(d.sort_by_in_degree)[:3] == ['cpu_tensor', 'shuffle_ratio', 'topo']
  assert _params(d.Dataset.__init__)[1:7] == ['graph', 'node_features', 'edge_features', 'node_labels', 'edge_dir',
                                              'node_split']
  for name, args in {
      'init_graph': ['edge_index', 'edge_ids', 'edge_weights', 'layout', 'graph_mode', 'directed', 'device'],
      'init_node_features': ['node_feature_data', 'id2idx', 'sort_func', 'split_ratio', 'device_group_list', 'device',
                             'with_gpu', 'dtype'],
      'init_edge_features': ['edge_feature_data', 'id2idx', 'split_ratio', 'device_group_list', 'device', 'with_gpu',
                             'dtype'],
      'init_node_labels': ['node_label_data'], 'init_node_split': ['node_split'],
      'random_node_split': ['num_val', 'num_test'],
      'load_vineyard': ['vineyard_id', 'vineyard_socket', 'edges', 'edge_weights', 'node_features', 'edge_features',
                        'node_labels']}.items():
    got = _params(getattr(d.Dataset, name))[1:]
    assert got[:len(args)] == args, (name, got)
  _has(d.Dataset, ['get_graph', 'get_node_types', 'get_edge_types', 'get_node_feature', 'get_edge_feature',
                   'get_node_label', 'share_ipc', 'from_ipc_handle'])
  _has(d, ['TableDataset', 'random_split', 'vineyard_to_csr', 'load_vertex_feature_from_vineyard',
           'load_edge_feature_from_vineyard', 'VineyardPartitionBook', 'VineyardGid2Lid', 'v6d_id_select', 'v6d_id_filter'])


def test_sampler_api():
  s = glt.sampler
  _has(s, ['NodeSamplerInput', 'EdgeSamplerInput', 'NegativeSampling', 'SamplerOutput', 'HeteroSamplerOutput',
           'NeighborOutput', 'SamplingType', 'SamplingConfig', 'NeighborSampler', 'RandomNegativeSampler', 'BaseSampler'])
  assert {'NODE', 'LINK', 'SUBGRAPH', 'RANDOM_WALK'} <= set(s.SamplingType.__members__)
  assert _params(s.NeighborSampler.__init__)[1:10] == ['graph', 'num_neighbors', 'device', 'with_edge', 'with_neg',
                                                       'with_weight', 'strategy', 'edge_dir', 'seed']
  _has(s.NeighborSampler, ['sample_one_hop', 'sample_from_nodes', 'sample_from_edges', 'sample_pyg_v1', 'subgraph',
                           'sample_prob', 'random_walk'])
  assert _params(s.RandomNegativeSampler.__init__)[1:4] == ['graph', 'mode', 'edge_dir']
  assert _params(s.RandomNegativeSampler.sample)[1:4] == ['req_num', 'trials_num', 'padding']
  assert _params(s.NegativeSampling.__init__)[1:4] == ['mode', 'amount', 'weight']
  for f in ('node', 'row', 'col', 'edge', 'batch', 'num_sampled_nodes', 'num_sampled_edges', 'device', 'metadata'):
    assert f in _params(s.SamplerOutput.__init__), f


def test_loader_api():
  ld = glt.loader
  _has(ld, ['NeighborLoader', 'LinkNeighborLoader', 'SubGraphLoader', 'NodeLoader', 'LinkLoader', 'to_data',
            'to_hetero_data', 'Data', 'HeteroData'])
  p = _params(ld.NeighborLoader.__init__)
  for a in ('data', 'num_neighbors', 'input_nodes', 'neighbor_sampler', 'batch_size', 'shuffle', 'drop_last', 'with_edge',
            'with_weight', 'strategy', 'device', 'as_pyg_v1', 'seed'):
    assert a in p, a
  p = _params(ld.LinkNeighborLoader.__init__)
  for a in ('data', 'num_neighbors', 'neighbor_sampler', 'edge_label_index', 'edge_label', 'neg_sampling', 'with_edge',
            'with_weight', 'batch_size', 'shuffle', 'drop_last', 'strategy', 'device', 'seed'):
    assert a in p, a
  assert _params(ld.SubGraphLoader.__init__)[1:4] == ['data', 'input_nodes', 'num_neighbors']


def test_partition_channel_utils_api():
  pt = glt.partition
  _has(pt, ['RandomPartitioner', 'FrequencyPartitioner', 'PartitionerBase', 'load_partition', 'cat_feature_cache',
            'build_partition_feature', 'PartitionBook', 'GLTPartitionBook', 'RangePartitionBook', 'save_meta', 'save_node_pb',
            'save_edge_pb', 'save_graph_partition', 'save_feature_partition'])
  assert _params(pt.PartitionerBase.partition)[1:3] == ['with_feature', 'graph_caching']
  for a in ('probs', 'cache_memory_budget', 'cache_ratio'):
    assert a in _params(pt.FrequencyPartitioner.__init__), a
  _has(glt.channel, ['ChannelBase', 'MpChannel', 'ShmChannel', 'RemoteReceivingChannel', 'SampleMessage', 'QueueTimeoutError'])
  _has(glt.utils, ['get_free_port', 'parse_size', 'save_ckpt', 'load_ckpt', 'id2idx', 'RandomSeedManager', 'share_memory',
                   'convert_to_tensor', 'coo_to_csr', 'coo_to_csc', 'ptr2ind', 'ind2ptr'])
  _has(glt.typing, ['NodeType', 'EdgeType', 'as_str', 'reverse_edge_type', 'Split'])


def test_distributed_api():
  d = glt.distributed
  _has(d, ['init_worker_group', 'get_context', 'DistRole', 'DistContext', 'init_rpc', 'shutdown_rpc', 'rpc_is_initialized',
           'barrier', 'all_gather', 'rpc_register', 'rpc_request', 'rpc_request_async', 'DistDataset', 'DistGraph',
           'DistFeature', 'DistNeighborSampler', 'DistLoader', 'DistNeighborLoader', 'DistLinkNeighborLoader',
           'DistSubGraphLoader', 'CollocatedDistSamplingWorkerOptions', 'MpDistSamplingWorkerOptions',
           'RemoteDistSamplingWorkerOptions', 'DistMpSamplingProducer', 'DistCollocatedSamplingProducer', 'DistServer',
           'init_server', 'wait_and_shutdown_server', 'init_client', 'shutdown_client', 'request_server',
           'async_request_server', 'DistRandomPartitioner', 'ConcurrentEventLoop', 'DistTableDataset'])
  p = _params(d.DistNeighborLoader.__init__)
  for a in ('data', 'num_neighbors', 'input_nodes', 'batch_size', 'shuffle', 'drop_last', 'with_edge', 'with_weight',
            'edge_dir', 'collect_features', 'to_device', 'random_seed', 'worker_options'):
    assert a in p, a
  _has(d.DistServer, ['get_dataset_meta', 'get_node_partition_id', 'get_node_feature', 'get_tensor_size', 'get_node_label',
                      'get_edge_index', 'get_edge_size', 'create_sampling_producer', 'destroy_sampling_producer',
                      'start_new_epoch_sampling', 'fetch_one_sampled_message'])
  _has(d.DistDataset, ['load', 'load_vineyard', 'random_node_split', 'share_ipc', 'from_ipc_handle', 'from_p2p'])
