"""Sampling producers: run DistNeighborSamplers in subprocesses (feeding a channel) or
collocated with the trainer.

Parity: reference python/distributed/dist_sampling_producer.py:41-365.  Differences: worker
exceptions are shipped to the consumer through a status queue and re-raised (the reference
only logs them and the epoch can hang); seed shuffling is seeded by
(seed, epoch) so an epoch can be replayed.
"""
import queue
import time
from enum import Enum
from typing import Optional, Union

import torch
import torch.multiprocessing as mp

from ..channel import ChannelBase
from ..sampler import EdgeSamplerInput, NodeSamplerInput, SamplingConfig, SamplingType
from ..utils.device import ensure_device
from .dist_context import get_context, init_worker_group
from .dist_dataset import DistDataset
from .dist_neighbor_sampler import DistNeighborSampler
from .dist_options import _BasicDistSamplingWorkerOptions
from .rpc import init_rpc, shutdown_rpc

MP_STATUS_CHECK_INTERVAL = 5.0


class MpCommand(Enum):
  SAMPLE_ALL = 0
  STOP = 1


def _batches(n: int, batch_size: int, drop_last: bool):
  for b in range(0, n, batch_size):
    if drop_last and b + batch_size > n:
      return
    yield slice(b, min(b + batch_size, n))


def _run_one(sampler: DistNeighborSampler, cfg: SamplingConfig, sampler_input, index):
  if cfg.sampling_type == SamplingType.NODE:
    return sampler.sample_from_nodes(sampler_input[index])
  if cfg.sampling_type == SamplingType.LINK:
    return sampler.sample_from_edges(sampler_input[index])
  if cfg.sampling_type == SamplingType.SUBGRAPH:
    return sampler.subgraph(sampler_input[index])
  raise NotImplementedError(cfg.sampling_type)


def _sampling_worker_loop(rank: int, data: DistDataset, sampler_input, unshuffled_index,
                          sampling_config: SamplingConfig, worker_options: _BasicDistSamplingWorkerOptions,
                          channel: ChannelBase, task_queue, status_queue, mp_barrier):
  """Body of one sampling subprocess."""
  dist_sampler = None
  graceful = False
  try:
    ctx = get_context()
    # the sampling subprocesses form their own RPC world
    group = (ctx.group_name if ctx is not None else 'sampler') + '_sampling'
    init_worker_group(worker_options.worker_world_size, worker_options.worker_ranks[rank], group)
    device = worker_options.worker_devices[rank]
    if torch.device(device).type == 'cuda':
      ensure_device(device)
    threads = worker_options.num_rpc_threads or min(data.num_partitions, 16)
    if data.num_partitions > 1 and getattr(data, 'data_plane', 'rpc') == 'rpc':
      init_rpc(worker_options.master_addr, worker_options.master_port, max(threads, 4),
               worker_options.rpc_timeout)
    dist_sampler = DistNeighborSampler(
      data, sampling_config.num_neighbors, sampling_config.with_edge, sampling_config.with_neg,
      sampling_config.with_weight, sampling_config.edge_dir, sampling_config.collect_features, channel,
      getattr(worker_options, 'use_all2all', False), worker_options.worker_concurrency, device,
      seed=sampling_config.seed)
    dist_sampler.start_loop()
    status_queue.put(('ready', rank, None))
    while True:
      try:
        cmd, args = task_queue.get(timeout=MP_STATUS_CHECK_INTERVAL)
      except queue.Empty:
        continue
      if cmd == MpCommand.STOP:
        # keep serving peers' remote requests until *every* sampling worker is done
        graceful = True
        break
      index, epoch = args
      n = 0
      for sl in _batches(index.numel(), sampling_config.batch_size, sampling_config.drop_last):
        _run_one(dist_sampler, sampling_config, sampler_input, index[sl])
        n += 1
      dist_sampler.wait_all()
      status_queue.put(('done', rank, n))
  except Exception as e:  # noqa: BLE001 -- surfaced to the consumer
    import traceback
    status_queue.put(('error', rank, f'{type(e).__name__}: {e}\n{traceback.format_exc()}'))
  finally:
    if dist_sampler is not None:
      try:
        dist_sampler.shutdown_loop()
      except Exception:  # noqa: BLE001
        pass
    try:
      shutdown_rpc(graceful=graceful)
    except Exception:  # noqa: BLE001
      pass


class DistMpSamplingProducer(object):
  """Spawns `num_workers` sampling subprocesses that push SampleMessages into `output_channel`."""

  def __init__(self, data: DistDataset, sampler_input: Union[NodeSamplerInput, EdgeSamplerInput],
               sampling_config: SamplingConfig, worker_options, output_channel: ChannelBase):
    self.data = data
    self.sampler_input = sampler_input.share_memory()
    self.input_len = len(self.sampler_input)
    self.sampling_config = sampling_config
    self.worker_options = worker_options
    self.worker_options._assign_worker_devices()
    ctx = get_context()
    if self.worker_options.worker_ranks is None and ctx is not None:
      self.worker_options._set_worker_ranks(ctx)
    self.num_workers = worker_options.num_workers
    self.output_channel = output_channel
    self._task_queues, self._workers = [], []
    self._status_queue = None
    self._epoch = 0
    self._shutdown = False
    self._pending_done = 0

  def init(self):
    mp_ctx = mp.get_context('spawn')
    self._status_queue = mp_ctx.Queue()
    for rank in range(self.num_workers):
      tq = mp_ctx.Queue(self.num_workers * 4)
      self._task_queues.append(tq)
      w = mp_ctx.Process(target=_sampling_worker_loop,
                         args=(rank, self.data, self.sampler_input, None, self.sampling_config,
                               self.worker_options, self.output_channel, tq, self._status_queue, None))
      w.daemon = True
      w.start()
      self._workers.append(w)
    # wait until every worker reports ready; fail fast on worker errors / deaths
    ready, t0 = 0, time.time()
    while ready < self.num_workers:
      try:
        kind, rank, info = self._status_queue.get(timeout=1.0)
      except queue.Empty:
        if any(not w.is_alive() for w in self._workers):
          raise RuntimeError('a sampling worker died during start-up')
        if time.time() - t0 > 600:
          raise TimeoutError('sampling workers failed to start within 600 s')
        continue
      if kind == 'error':
        raise RuntimeError(f'sampling worker {rank} failed during start-up:\n{info}')
      if kind == 'ready':
        ready += 1

  def _check_errors(self, block: bool = False):
    while True:
      try:
        kind, rank, info = self._status_queue.get(timeout=0.1) if block else self._status_queue.get_nowait()
      except queue.Empty:
        return
      if kind == 'error':
        raise RuntimeError(f'sampling worker {rank} failed:\n{info}')
      if kind == 'done':
        self._pending_done -= 1

  def shutdown(self):
    if self._shutdown:
      return
    self._shutdown = True
    for q in self._task_queues:
      try:
        q.put((MpCommand.STOP, None))
      except Exception:  # noqa: BLE001
        pass
    for w in self._workers:
      w.join(timeout=120)   # graceful RPC shutdown waits for the slowest peer worker
    for w in self._workers:
      if w.is_alive():
        w.terminate()

  def __del__(self):
    try:
      self.shutdown()
    except Exception:  # noqa: BLE001
      pass

  def produce_all(self):
    """Start one epoch: split the (optionally shuffled) index range over the workers."""
    cfg = self.sampling_config
    if cfg.shuffle:
      g = torch.Generator()
      g.manual_seed((cfg.seed or 0) + self._epoch)
      index = torch.randperm(self.input_len, generator=g)
    else:
      index = torch.arange(self.input_len)
    self._epoch += 1
    # whole batches per worker so that the number of messages is deterministic
    n_batches = (self.input_len // cfg.batch_size) if cfg.drop_last else \
        (self.input_len + cfg.batch_size - 1) // cfg.batch_size
    per = (n_batches + self.num_workers - 1) // self.num_workers
    self._pending_done = 0
    for r, q in enumerate(self._task_queues):
      lo, hi = r * per * cfg.batch_size, min((r + 1) * per * cfg.batch_size, self.input_len)
      if cfg.drop_last:
        hi = min(hi, n_batches * cfg.batch_size)
      part = index[lo:hi] if lo < hi else index[:0]
      q.put((MpCommand.SAMPLE_ALL, (part, self._epoch)))
      self._pending_done += 1
    return n_batches

  def is_all_sampling_completed(self) -> bool:
    """Every worker reported the end of its share of the epoch (messages may still sit in the channel)."""
    self._check_errors()
    return self._pending_done <= 0

  def is_all_sampling_completed_and_consumed(self) -> bool:
    return self.is_all_sampling_completed() and self.output_channel.empty()

  def check_errors(self):
    self._check_errors()


class DistCollocatedSamplingProducer(object):
  """Sampler living in the trainer process; `sample()` returns one message synchronously."""

  def __init__(self, data: DistDataset, sampler_input, sampling_config: SamplingConfig, worker_options, device):
    self.data = data
    self.sampler_input = sampler_input
    self.sampling_config = sampling_config
    self.worker_options = worker_options
    self.device = device
    self._epoch = 0
    self._iter = None

  def init(self):
    cfg = self.sampling_config
    self._sampler = DistNeighborSampler(self.data, cfg.num_neighbors, cfg.with_edge, cfg.with_neg, cfg.with_weight,
                                        cfg.edge_dir, cfg.collect_features, None,
                                        getattr(self.worker_options, 'use_all2all', False), 1, self.device,
                                        seed=cfg.seed)
    self._sampler.start_loop()

  def shutdown(self):
    if getattr(self, '_sampler', None) is not None:
      self._sampler.shutdown_loop()

  def reset(self):
    cfg = self.sampling_config
    n = len(self.sampler_input)
    if cfg.shuffle:
      g = torch.Generator()
      g.manual_seed((cfg.seed or 0) + self._epoch)
      index = torch.randperm(n, generator=g)
    else:
      index = torch.arange(n)
    self._epoch += 1
    self._iter = iter([index[sl] for sl in _batches(n, cfg.batch_size, cfg.drop_last)])

  def sample(self):
    index = next(self._iter)
    return _run_one(self._sampler, self.sampling_config, self.sampler_input, index)
