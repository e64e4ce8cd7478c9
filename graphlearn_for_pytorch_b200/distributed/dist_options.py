"""Sampling worker options (parity: reference python/distributed/dist_options.py:25-298)."""
import os
from typing import List, Optional, Union

import torch

from ..utils.units import parse_size
from .dist_context import assign_server_by_order, get_context


class _BasicDistSamplingWorkerOptions(object):
  """Options shared by every deployment mode.

  num_workers: sampling workers to launch.
  worker_devices: device per worker (auto-assigned round-robin when None).
  worker_concurrency: in-flight batches per worker (clamped to 1..32).
  master_addr / master_port: rendezvous of the *sampling* RPC group (defaults:
    $MASTER_ADDR, $MASTER_PORT + 1).
  num_rpc_threads / rpc_timeout: RPC agent tuning.
  """

  def __init__(self, num_workers: int = 1, worker_devices=None, worker_concurrency: int = 1,
               master_addr: Optional[str] = None, master_port: Optional[Union[str, int]] = None,
               num_rpc_threads: Optional[int] = None, rpc_timeout: float = 180):
    self.num_workers = num_workers
    self.worker_world_size = None
    self.worker_ranks = None
    if worker_devices is None:
      self.worker_devices = None
    elif isinstance(worker_devices, (list, tuple)):
      assert len(worker_devices) == num_workers
      self.worker_devices = list(worker_devices)
    else:
      self.worker_devices = [worker_devices] * num_workers
    self.worker_concurrency = max(1, min(int(worker_concurrency), 32))
    self.master_addr = str(master_addr) if master_addr is not None else os.environ.get('MASTER_ADDR')
    if self.master_addr is None:
      raise ValueError('missing master address for the sampling RPC group (set MASTER_ADDR)')
    if master_port is not None:
      self.master_port = int(master_port)
    elif os.environ.get('MASTER_PORT') is not None:
      self.master_port = int(os.environ['MASTER_PORT']) + 1
    else:
      raise ValueError('missing master port for the sampling RPC group (set MASTER_PORT)')
    self.num_rpc_threads = num_rpc_threads
    if num_rpc_threads is not None:
      assert num_rpc_threads > 0
    self.rpc_timeout = rpc_timeout

  def _set_worker_ranks(self, current_ctx):
    self.worker_world_size = current_ctx.world_size * self.num_workers
    self.worker_ranks = [current_ctx.rank * self.num_workers + i for i in range(self.num_workers)]

  def _assign_worker_devices(self):
    if self.worker_devices is not None:
      return
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    self.worker_devices = [torch.device('cuda', i % n) if n > 0 else torch.device('cpu')
                           for i in range(self.num_workers)]


class CollocatedDistSamplingWorkerOptions(_BasicDistSamplingWorkerOptions):
  """Sample synchronously inside the training process (no channel).  With
  `use_all2all=True` remote features are exchanged with collectives instead of RPC."""

  def __init__(self, master_addr=None, master_port=None, num_rpc_threads=None, rpc_timeout: float = 180,
               use_all2all: bool = False):
    super().__init__(1, None, 1, master_addr, master_port, num_rpc_threads, rpc_timeout)
    self.use_all2all = use_all2all


class MpDistSamplingWorkerOptions(_BasicDistSamplingWorkerOptions):
  """Sample in spawned subprocesses feeding a shared-memory channel.

  channel_capacity: messages the channel may hold (default num_workers * concurrency).
  channel_size: shm bytes (default num_workers * 64MB).  pin_memory: page-lock the ring.
  """

  def __init__(self, num_workers: int = 1, worker_devices=None, worker_concurrency: int = 4,
               master_addr=None, master_port=None, num_rpc_threads=None, rpc_timeout: float = 180,
               channel_size: Optional[Union[int, str]] = None, pin_memory: bool = False,
               use_all2all: bool = False):
    super().__init__(num_workers, worker_devices, worker_concurrency, master_addr, master_port,
                     num_rpc_threads, rpc_timeout)
    self.channel_capacity = self.num_workers * self.worker_concurrency
    self.channel_size = parse_size(channel_size) if channel_size is not None else self.num_workers * (64 << 20)
    self.pin_memory = pin_memory
    self.use_all2all = use_all2all


class RemoteDistSamplingWorkerOptions(_BasicDistSamplingWorkerOptions):
  """Sample on remote servers (server-client mode); the client pulls messages.

  server_rank: server(s) to use (default: assignment by order).  buffer_size: server-side
  shm buffer.  prefetch_size: outstanding fetch RPCs per server.  worker_key: identifies a
  producer so that loaders with the same key share it.
  """

  def __init__(self, server_rank: Optional[Union[int, List[int]]] = None, num_workers: int = 1,
               worker_devices=None, worker_concurrency: int = 4, master_addr=None, master_port=None,
               num_rpc_threads=None, rpc_timeout: float = 180, buffer_size: Optional[Union[int, str]] = None,
               prefetch_size: int = 4, worker_key: Optional[str] = None, glt_graph=None,
               workload_type: Optional[str] = None, use_all2all: bool = False):
    # GraphScope hands over a handle that carries the rendezvous address and one loader port per workload
    # ('train' | 'validate' | 'test'), reference dist_options.py:257-272
    if glt_graph is not None:
      if workload_type not in ('train', 'validate', 'test'):
        raise ValueError(f"'{self.__class__.__name__}': workload_type must be 'train', 'validate' or 'test' "
                         f"when glt_graph is given")
      master_addr = glt_graph.master_addr
      master_port = {'train': 'train_loader_master_port', 'validate': 'val_loader_master_port',
                     'test': 'test_loader_master_port'}[workload_type]
      master_port = getattr(glt_graph, master_port)
    super().__init__(num_workers, worker_devices, worker_concurrency, master_addr, master_port,
                     num_rpc_threads, rpc_timeout)
    if server_rank is not None:
      self.server_rank = server_rank
    else:
      ctx = get_context()
      self.server_rank = assign_server_by_order() if ctx is not None and ctx.is_client() else 0
    self.buffer_capacity = self.num_workers * self.worker_concurrency
    self.buffer_size = parse_size(buffer_size) if buffer_size is not None else f'{self.num_workers * 64}MB'
    self.prefetch_size = prefetch_size
    if self.prefetch_size > self.buffer_capacity:
      raise ValueError(f'prefetch_size {prefetch_size} exceeds the server buffer capacity {self.buffer_capacity}')
    self.worker_key = worker_key
    self.use_all2all = use_all2all


AllDistSamplingWorkerOptions = Union[CollocatedDistSamplingWorkerOptions, MpDistSamplingWorkerOptions,
                                     RemoteDistSamplingWorkerOptions]
