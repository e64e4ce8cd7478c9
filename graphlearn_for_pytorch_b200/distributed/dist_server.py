"""Sampling server of the server-client deployment mode
(parity: reference python/distributed/dist_server.py:39-296)."""
import logging
import threading
import time
from typing import Dict, Optional, Union

import torch

from ..channel import QueueTimeoutError, ShmChannel
from ..sampler import EdgeSamplerInput, NodeSamplerInput, RemoteSamplerInput, SamplingConfig
from .dist_context import _set_server_context, get_context
from .dist_dataset import DistDataset
from .dist_options import RemoteDistSamplingWorkerOptions
from .dist_sampling_producer import DistMpSamplingProducer
from .rpc import barrier, init_rpc, shutdown_rpc

SERVER_EXIT_STATUS_CHECK_INTERVAL = 5.0


class DistServer(object):
  """Holds a dataset partition, serves metadata / feature / label queries and runs sampling
  producers on behalf of clients."""

  def __init__(self, dataset: DistDataset):
    self.dataset = dataset
    self._lock = threading.RLock()
    self._exit = False
    self._cur_producer_idx = 0
    self._producer_pool: Dict[int, DistMpSamplingProducer] = {}
    self._buffer_pool: Dict[int, ShmChannel] = {}
    self._key_to_producer: Dict[str, int] = {}
    self._epoch: Dict[int, int] = {}
    self._ready: Dict[int, threading.Event] = {}

  def shutdown(self):
    for pid in list(self._producer_pool.keys()):
      self.destroy_sampling_producer(pid)

  def wait_for_exit(self):
    while not self._exit:
      time.sleep(SERVER_EXIT_STATUS_CHECK_INTERVAL)

  def exit(self):
    self._exit = True
    return True

  # ---- dataset queries (PyG remote-backend surface)
  def get_dataset_meta(self):
    return (self.dataset.num_partitions, self.dataset.partition_idx, self.dataset.get_node_types(),
            self.dataset.get_edge_types())

  def get_node_partition_id(self, node_type, index: torch.Tensor):
    pb = self.dataset.node_pb[node_type] if isinstance(self.dataset.node_pb, dict) else self.dataset.node_pb
    return pb[index]

  def get_node_feature(self, node_type, index: torch.Tensor):
    feat = self.dataset.get_node_feature(node_type)
    return feat.cpu_get(index) if feat is not None else None

  def get_tensor_size(self, node_type):
    feat = self.dataset.get_node_feature(node_type)
    return torch.Size(feat.shape) if feat is not None else None

  def get_node_label(self, node_type, index: torch.Tensor):
    lab = self.dataset.get_node_label(node_type)
    return lab[index] if lab is not None else None

  def get_edge_index(self, edge_type, layout: str = 'coo'):
    g = self.dataset.get_graph(edge_type)
    row, col, _, _ = g.topo.to_coo()
    return row, col

  def get_edge_size(self, edge_type, layout: str = 'coo'):
    g = self.dataset.get_graph(edge_type)
    return g.row_count, g.col_count

  # ---- sampling producers
  def create_sampling_producer(self, sampler_input, sampling_config: SamplingConfig,
                               worker_options: RemoteDistSamplingWorkerOptions) -> int:
    if isinstance(sampler_input, RemoteSamplerInput):
      sampler_input = sampler_input.to_local_sampler_input(dataset=self.dataset)
    # NB: the lock must NOT be held while the sampling workers start.  Their RPC group spans all servers, so
    # server A blocks in init() until server B has started the workers of the SAME client; if B is meanwhile
    # (under its lock) starting another client's workers, which in turn wait for A, the two servers deadlock.
    with self._lock:
      key = worker_options.worker_key
      if key is not None and key in self._key_to_producer:
        pid = self._key_to_producer[key]
        ready = self._ready.get(pid)
      else:
        ready = None
        pid = self._cur_producer_idx
        self._cur_producer_idx += 1
        buf = ShmChannel(worker_options.buffer_capacity, worker_options.buffer_size)
        ctx = get_context()
        worker_options._set_worker_ranks(ctx)
        prod = DistMpSamplingProducer(self.dataset, sampler_input, sampling_config, worker_options, buf)
        self._producer_pool[pid] = prod
        self._buffer_pool[pid] = buf
        self._epoch[pid] = -1
        self._ready[pid] = threading.Event()
        if key is not None:
          self._key_to_producer[key] = pid
    if ready is not None:          # another request is (or was) creating this producer: wait until it is usable
      ready.wait()
      return pid
    try:
      prod.init()
    finally:
      self._ready[pid].set()
    return pid

  def destroy_sampling_producer(self, producer_id: int):
    with self._lock:
      prod = self._producer_pool.pop(producer_id, None)
      self._buffer_pool.pop(producer_id, None)
      self._ready.pop(producer_id, None)
      for k, v in list(self._key_to_producer.items()):
        if v == producer_id:
          self._key_to_producer.pop(k)
    if prod is not None:
      prod.shutdown()
    return True

  def start_new_epoch_sampling(self, producer_id: int, epoch: int = 0):
    with self._lock:
      if self._epoch.get(producer_id, -1) < epoch:
        self._epoch[producer_id] = epoch
        self._producer_pool[producer_id].produce_all()
    return True

  def fetch_one_sampled_message(self, producer_id: int):
    """-> (message | None, end_of_epoch)."""
    prod = self._producer_pool.get(producer_id)
    buf = self._buffer_pool.get(producer_id)
    if prod is None:
      return None, True
    while True:
      try:
        msg = buf.recv(timeout_ms=500)
        return {k: v.clone() for k, v in msg.items()}, False
      except QueueTimeoutError:
        if prod.is_all_sampling_completed_and_consumed():
          return None, True


_dist_server: Optional[DistServer] = None


def get_server() -> Optional[DistServer]:
  return _dist_server


def init_server(num_servers: int, server_rank: int, dataset: DistDataset, master_addr: str, master_port: int,
                num_clients: int = 0, num_rpc_threads: int = 16, request_timeout: int = 180,
                server_group_name: Optional[str] = None, is_dynamic: bool = False):
  """Declare this process as sampling server `server_rank` and join the RPC world."""
  global _dist_server
  _set_server_context(num_servers, server_rank, server_group_name, num_clients)
  _dist_server = DistServer(dataset)
  init_rpc(master_addr, master_port, num_rpc_threads, request_timeout, is_dynamic=is_dynamic)
  return _dist_server


def wait_and_shutdown_server():
  """Block until a client calls DistServer.exit, then tear everything down."""
  global _dist_server
  ctx = get_context()
  if ctx is None or not ctx.is_server():
    raise RuntimeError('wait_and_shutdown_server() must be called from a server process')
  _dist_server.wait_for_exit()
  _dist_server.shutdown()
  _dist_server = None
  barrier()
  shutdown_rpc()


def _call_func_on_server(func, *args, **kwargs):
  """RPC trampoline: run `func(server, ...)` inside the server process."""
  if not callable(func):
    logging.warning('non-callable object received by the server: %r', func)
    return None
  srv = get_server()
  if srv is None:
    raise RuntimeError('this process is not an initialised DistServer')
  return func(srv, *args, **kwargs)
