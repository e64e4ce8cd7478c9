"""Client-side pull channel for the server-client deployment mode
(parity: reference python/channel/remote_channel.py:24-131): keeps `prefetch_size`
fetch RPCs in flight per sampling server and stops when every server reported the end
of the epoch.  Errors raised on the server side are re-raised on the consumer instead
of being logged and dropped."""
import queue
import threading
from typing import List, Union

from .base import ChannelBase, SampleMessage


class RemoteReceivingChannel(ChannelBase):
  """Client side of the server-client mode: keeps `prefetch_size` asynchronous `fetch_one_sampled_message` requests
  in flight per server and hands the messages out in arrival order; end of epoch is signalled by the servers
  (reference: python/channel/remote_channel.py:24-131)."""
  def __init__(self, server_rank: Union[int, List[int]], producer_id: Union[int, List[int]],
               prefetch_size: int = 4):
    self.server_rank_list = server_rank if isinstance(server_rank, list) else [server_rank]
    self.producer_id_list = producer_id if isinstance(producer_id, list) else [producer_id]
    assert len(self.server_rank_list) == len(self.producer_id_list)
    self.prefetch_size = prefetch_size
    self.num_expected = -1  # unknown: rely on end-of-epoch flags
    self._lock = threading.RLock()
    self.reset()

  def reset(self):
    with self._lock:
      self._queue = queue.Queue()
      self._inflight = {r: 0 for r in self.server_rank_list}
      self._ended = {r: False for r in self.server_rank_list}
      self._global_end = False

  def send(self, msg: SampleMessage, **kwargs):
    raise RuntimeError('RemoteReceivingChannel is receive-only')

  def _request_more(self):
    from ..distributed import dist_client, dist_server  # lazy: breaks the import cycle
    with self._lock:
      for rank, pid in zip(self.server_rank_list, self.producer_id_list):
        while not self._ended[rank] and self._inflight[rank] < self.prefetch_size:
          fut = dist_client.async_request_server(rank, dist_server.DistServer.fetch_one_sampled_message, pid)
          self._inflight[rank] += 1
          fut.add_done_callback(lambda f, r=rank: self._on_done(f, r))

  def _on_done(self, fut, rank):
    try:
      msg, end = fut.wait()
    except Exception as e:  # surface failures to the consumer
      self._queue.put(e)
      with self._lock:
        self._inflight[rank] -= 1
        self._ended[rank] = True
      return
    with self._lock:
      self._inflight[rank] -= 1
      if end:
        self._ended[rank] = True
    self._queue.put((msg, end, rank))

  def recv(self, **kwargs) -> SampleMessage:
    while True:
      with self._lock:
        done = all(self._ended.values()) and all(v == 0 for v in self._inflight.values())
      if done and self._queue.empty():
        raise StopIteration
      self._request_more()
      try:
        item = self._queue.get(timeout=0.5)
      except queue.Empty:
        continue
      if isinstance(item, Exception):
        raise item
      msg, end, _ = item
      if msg is not None:
        return msg
