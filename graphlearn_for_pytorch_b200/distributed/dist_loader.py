"""DistLoader: the base of the distributed loaders (`dist_neighbor_loader.py`, `dist_link_neighbor_loader.py`,
`dist_subgraph_loader.py`).

Parity: reference python/distributed/dist_loader.py:46-451.  A loader owns (a) a sampling
producer chosen by the worker options (collocated / multiprocess / remote server) and
(b) the message -> Data/HeteroData collation.
"""
from typing import Optional

import torch

from ..channel import QueueTimeoutError, RemoteReceivingChannel, SampleMessage, ShmChannel
from ..loader.transform import to_data, to_hetero_data
from ..sampler import HeteroSamplerOutput, SamplerOutput, SamplingConfig
from ..typing import from_str
from ..utils.exit_status import is_python_exiting
from .dist_context import get_context
from .dist_dataset import DistDataset
from .dist_options import (AllDistSamplingWorkerOptions, CollocatedDistSamplingWorkerOptions,
                           MpDistSamplingWorkerOptions, RemoteDistSamplingWorkerOptions)
from .dist_sampling_producer import DistCollocatedSamplingProducer, DistMpSamplingProducer
from .rpc import init_rpc, rpc_is_initialized


def message_to_sampler_output(msg: SampleMessage, device, edge_dir: str = 'out'):
  """Inverse of DistNeighborSampler._colloate_fn: -> (output, x, y, edge_attr) or hetero dicts."""
  is_hetero = bool(int(msg['#IS_HETERO'][0]))
  to = (lambda t: t.to(device, non_blocking=True)) if device is not None else (lambda t: t)
  meta = {k[len('#META.'):]: to(v) for k, v in msg.items() if k.startswith('#META.')}
  input_type = None
  if 'input_type' in meta:
    input_type = from_str(bytes(meta.pop('input_type').cpu().tolist()).decode())
  if not is_hetero:
    md = meta if meta and 'mapping' not in meta else (meta.get('mapping') if meta else None)
    out = SamplerOutput(
      node=to(msg['ids']), row=to(msg['rows']), col=to(msg['cols']),
      edge=to(msg['eids']) if 'eids' in msg else None,
      batch=to(msg['batch']) if 'batch' in msg else None,
      # per-hop counts stay int64 tensors, like the reference's loader hands them out (dist_loader.py:428-447)
      num_sampled_nodes=msg['num_sampled_nodes'].to(torch.int64).cpu() if 'num_sampled_nodes' in msg else None,
      num_sampled_edges=msg['num_sampled_edges'].to(torch.int64).cpu() if 'num_sampled_edges' in msg else None,
      device=device, metadata=md)
    return out, (to(msg['nfeats']) if 'nfeats' in msg else None), \
        (to(msg['nlabels']) if 'nlabels' in msg else None), (to(msg['efeats']) if 'efeats' in msg else None)
  node, batch, nsn, x, y = {}, {}, {}, {}, {}
  row, col, edge, nse, ea = {}, {}, {}, {}, {}
  for k, v in msg.items():
    if k.startswith('#') or '.' not in k:
      continue
    t, attr = k.rsplit('.', 1)
    t = from_str(t)
    if attr == 'ids':
      node[t] = to(v)
    elif attr == 'batch':
      batch[t] = to(v)
    elif attr == 'num_sampled_nodes':
      nsn[t] = v.to(torch.int64).cpu()
    elif attr == 'nfeats':
      x[t] = to(v)
    elif attr == 'nlabels':
      y[t] = to(v)
    elif attr == 'rows':
      row[t] = to(v)
    elif attr == 'cols':
      col[t] = to(v)
    elif attr == 'eids':
      edge[t] = to(v)
    elif attr == 'num_sampled_edges':
      nse[t] = v.to(torch.int64).cpu()
    elif attr == 'efeats':
      ea[t] = to(v)
  out = HeteroSamplerOutput(node=node, row=row, col=col, edge=edge or None, batch=batch or None,
                            num_sampled_nodes=nsn, num_sampled_edges=nse, edge_types=list(row.keys()),
                            input_type=input_type, device=device, metadata=meta or None)
  return out, (x or None), (y or None), (ea or None)


class DistLoader(object):
  """Args:
    data: DistDataset (None in remote/client mode).
    input_data: NodeSamplerInput / EdgeSamplerInput / RemoteSamplerInput.
    sampling_config: SamplingConfig.
    to_device: device of the yielded batches.
    worker_options: Collocated / Mp / Remote options (default: collocated).
  """

  def __init__(self, data: Optional[DistDataset], input_data, sampling_config: SamplingConfig,
               to_device: Optional[torch.device] = None,
               worker_options: Optional[AllDistSamplingWorkerOptions] = None):
    self.data = data
    self.input_data = input_data
    self.sampling_config = sampling_config
    self.sampling_type = sampling_config.sampling_type
    self.batch_size = sampling_config.batch_size
    self.drop_last = sampling_config.drop_last
    self.edge_dir = sampling_config.edge_dir
    self.to_device = torch.device(to_device) if to_device is not None else \
        (torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu'))
    self.worker_options = worker_options or CollocatedDistSamplingWorkerOptions(
      master_addr='127.0.0.1', master_port=29400)
    self._is_collocated = isinstance(self.worker_options, CollocatedDistSamplingWorkerOptions)
    self._is_mp = isinstance(self.worker_options, MpDistSamplingWorkerOptions)
    self._is_remote = isinstance(self.worker_options, RemoteDistSamplingWorkerOptions)
    self._epoch = 0
    self._shutdowned = False
    self._num_recv = 0
    self._num_expected = 0
    self._channel = None
    self._producer = None
    ctx = get_context()

    if self._is_remote:
      from . import dist_client, dist_server
      self._server_ranks = self.worker_options.server_rank if isinstance(self.worker_options.server_rank, list) \
          else [self.worker_options.server_rank]
      self._input_type = getattr(input_data, 'input_type', None) if not isinstance(input_data, list) \
          else getattr(input_data[0], 'input_type', None)
      inputs = input_data if isinstance(input_data, list) else [input_data] * len(self._server_ranks)
      (self.num_data_partitions, self.data_partition_idx, self._node_types, self._edge_types) = \
          dist_client.request_server(self._server_ranks[0], dist_server.DistServer.get_dataset_meta)
      # the sampling workers of all servers form one RPC world: create them concurrently,
      # a sequential (blocking) creation would deadlock in their rendezvous
      futs = [dist_client.async_request_server(srv, dist_server.DistServer.create_sampling_producer, inp,
                                               self.sampling_config, self.worker_options)
              for srv, inp in zip(self._server_ranks, inputs)]
      self._producer_ids = [f.wait() for f in futs]
      self._channel = RemoteReceivingChannel(self._server_ranks, self._producer_ids,
                                             self.worker_options.prefetch_size)
      return

    assert data is not None, 'a local DistDataset is required in worker mode'
    self.num_data_partitions = data.num_partitions
    self.data_partition_idx = data.partition_idx
    self._node_types = data.get_node_types()
    self._edge_types = data.get_edge_types()
    self._input_type = getattr(input_data, 'input_type', None)
    self._input_len = len(input_data)
    n_batches = (self._input_len // self.batch_size) if self.drop_last else \
        (self._input_len + self.batch_size - 1) // self.batch_size
    self._num_expected = n_batches
    if ctx is None:
      raise RuntimeError('init_worker_group() must be called before creating a DistLoader')
    if self._is_collocated:
      needs_rpc = data.num_partitions > 1 and getattr(data, 'data_plane', 'rpc') == 'rpc'
      if needs_rpc and not rpc_is_initialized():
        init_rpc(self.worker_options.master_addr, self.worker_options.master_port,
                 self.worker_options.num_rpc_threads or 16, self.worker_options.rpc_timeout)
      dev = self.to_device if data.get_graph() is None or not isinstance(data.graph, dict) else self.to_device
      self._producer = DistCollocatedSamplingProducer(data, input_data, sampling_config, self.worker_options, dev)
      self._producer.init()
    else:
      self.worker_options._set_worker_ranks(ctx)
      self._channel = ShmChannel(self.worker_options.channel_capacity, self.worker_options.channel_size)
      if self.worker_options.pin_memory:
        self._channel.pin_memory()
      self._producer = DistMpSamplingProducer(data, input_data, sampling_config, self.worker_options,
                                              self._channel)
      self._producer.init()

  # ------------------------------------------------------------------ lifecycle
  def __del__(self):
    if is_python_exiting():
      return
    self.shutdown()

  def shutdown(self):
    if self._shutdowned:
      return
    self._shutdowned = True
    if self._is_remote:
      try:
        from . import dist_client, dist_server
        futs = [dist_client.async_request_server(srv, dist_server.DistServer.destroy_sampling_producer, pid)
                for srv, pid in zip(self._server_ranks, self._producer_ids)]
        for f in futs:
          f.wait()
      except Exception:  # noqa: BLE001
        pass
    elif self._producer is not None:
      self._producer.shutdown()

  def __len__(self):
    return self._num_expected

  def __iter__(self):
    self._num_recv = 0
    if self._is_collocated:
      self._producer.reset()
    elif self._is_mp:
      self._num_expected = self._producer.produce_all()
    else:
      from . import dist_client, dist_server
      for srv, pid in zip(self._server_ranks, self._producer_ids):
        dist_client.request_server(srv, dist_server.DistServer.start_new_epoch_sampling, pid, self._epoch)
      self._channel.reset()
    self._epoch += 1
    return self

  def __next__(self):
    if self._is_remote:
      msg = self._channel.recv()      # raises StopIteration when every server is done
    elif self._is_collocated:
      msg = self._producer.sample()   # raises StopIteration at the end of the epoch
    else:
      if self._num_recv >= self._num_expected:
        raise StopIteration
      while True:
        try:
          msg = self._channel.recv(timeout_ms=2000)
          break
        except QueueTimeoutError:
          self._producer.check_errors()   # a dead worker becomes an exception, not a hang
    self._num_recv += 1
    return self._collate_fn(msg)

  # ------------------------------------------------------------------ collation
  def _collate_fn(self, msg: SampleMessage):
    out, x, y, ea = message_to_sampler_output(msg, self.to_device, self.edge_dir)
    if isinstance(out, HeteroSamplerOutput):
      return to_hetero_data(out, batch_label_dict=y, node_feat_dict=x, edge_feat_dict=ea, edge_dir=self.edge_dir)
    return to_data(out, batch_labels=y, node_feats=x, edge_feats=ea)


def _loader_repr(self) -> str:
  return f'{self.__class__.__name__}()'


DistLoader.__repr__ = _loader_repr


def __getattr__(name):
  # the three flavours lived in this module before they moved next to their reference counterparts
  import importlib
  home = {'DistNeighborLoader': 'dist_neighbor_loader', 'DistLinkNeighborLoader': 'dist_link_neighbor_loader',
          'DistSubGraphLoader': 'dist_subgraph_loader'}.get(name)
  if home is None:
    raise AttributeError(name)
  return getattr(importlib.import_module(f'{__package__}.{home}'), name)
