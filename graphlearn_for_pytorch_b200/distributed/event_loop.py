"""asyncio event loop on a daemon thread with bounded concurrency
(parity: reference python/distributed/event_loop.py:24-102).  Unlike the reference, an
exception inside a task is *stored and re-raised* to the caller of wait_all()/run_task()
instead of being logged and dropped."""
import asyncio
import threading
from typing import Callable, Optional

import torch


def wrap_torch_future(f: torch.futures.Future) -> asyncio.futures.Future:
  """torch Future -> awaitable asyncio Future bound to the running loop."""
  loop = asyncio.get_event_loop()
  aio = loop.create_future()

  def on_done(*_):
    try:
      res = f.value()
    except Exception as e:  # noqa: BLE001
      loop.call_soon_threadsafe(aio.set_exception, e)
    else:
      loop.call_soon_threadsafe(aio.set_result, res)
  f.add_done_callback(on_done)
  return aio


class ConcurrentEventLoop(object):
  """An asyncio loop on its own thread with a concurrency bound: `add_task(coro, callback)` schedules, `run_task`
  blocks for the result, `wait_all` drains.  Lets one sampler overlap the remote hops of several batches; task
  exceptions are re-raised to the caller instead of being logged and dropped (reference:
  python/distributed/event_loop.py:37-113)."""
  def __init__(self, concurrency: int):
    self._concurrency = concurrency
    self._sem = threading.BoundedSemaphore(concurrency)
    self._loop = asyncio.new_event_loop()
    self._thread = threading.Thread(target=self._run, daemon=True)
    self._errors = []
    self._lock = threading.Lock()

  def _run(self):
    asyncio.set_event_loop(self._loop)
    self._loop.run_forever()

  def start_loop(self):
    if not self._thread.is_alive():
      self._thread.start()

  def shutdown_loop(self):
    self.wait_all(raise_errors=False)
    if self._loop.is_running():
      self._loop.call_soon_threadsafe(self._loop.stop)
      self._thread.join(timeout=5)
    if not self._loop.is_running() and not self._loop.is_closed():
      self._loop.close()                       # releases the selector + self-pipe (else a ResourceWarning at exit)

  def wait_all(self, raise_errors: bool = True):
    """Block until every submitted task finished; re-raise the first task error."""
    for _ in range(self._concurrency):
      self._sem.acquire()
    for _ in range(self._concurrency):
      self._sem.release()
    if raise_errors:
      with self._lock:
        if self._errors:
          e = self._errors[0]
          self._errors = []
          raise e

  def add_task(self, coro, callback: Optional[Callable] = None):
    """Schedule `coro`; blocks while `concurrency` tasks are already in flight."""
    self._sem.acquire()

    def on_done(f: asyncio.futures.Future):
      try:
        res = f.result()
        if callback is not None:
          callback(res)
      except Exception as e:  # noqa: BLE001
        with self._lock:
          self._errors.append(e)
      finally:
        self._sem.release()
    fut = asyncio.run_coroutine_threadsafe(coro, self._loop)
    fut.add_done_callback(on_done)

  def run_task(self, coro):
    """Run `coro` to completion and return its result (exceptions propagate)."""
    with self._sem:
      return asyncio.run_coroutine_threadsafe(coro, self._loop).result()
