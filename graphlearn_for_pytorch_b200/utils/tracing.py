"""Tracing / timing / structured logging.

The reference has no profiling hooks at all and logs MLPerf
events only from its IGBH example (examples/igbh/mlperf_logging_utils.py:6-33).
Here: NVTX ranges around every pipeline stage, CUDA-event device timers (the
only clock used for reported numbers) and a tiny mllog-compatible event logger.
"""
import contextlib
import json
import time
from typing import Dict, List, Optional

import torch

_NVTX = torch.cuda.is_available()


@contextlib.contextmanager
def nvtx_range(name: str):
  if _NVTX:
    torch.cuda.nvtx.range_push(name)
  try:
    yield
  finally:
    if _NVTX:
      torch.cuda.nvtx.range_pop()


class DeviceTimer:
  """CUDA-event timer on the current stream; falls back to perf_counter on CPU."""

  def __init__(self):
    self.cuda = torch.cuda.is_available()
    self.records: Dict[str, List[float]] = {}

  @contextlib.contextmanager
  def measure(self, name: str):
    if self.cuda:
      s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      s.record()
      yield
      e.record()
      e.synchronize()
      ms = s.elapsed_time(e)
    else:
      t0 = time.perf_counter()
      yield
      ms = (time.perf_counter() - t0) * 1e3
    self.records.setdefault(name, []).append(ms)

  def summary(self) -> Dict[str, float]:
    return {k: sum(v) / len(v) for k, v in self.records.items() if v}


class EventLogger:
  """MLPerf-mllog style ':::MLLOG {json}' lines (INIT_START, RUN_START, EVAL_ACCURACY, ...)."""

  def __init__(self, path: Optional[str] = None, rank: int = 0):
    self.path, self.rank = path, rank

  def event(self, key: str, value=None, metadata: Optional[dict] = None):
    if self.rank != 0:
      return
    line = ':::MLLOG ' + json.dumps({'time_ms': int(time.time() * 1e3), 'key': key,
                                     'value': value, 'metadata': metadata or {}})
    if self.path:
      with open(self.path, 'a') as f:
        f.write(line + '\n')
    else:
      print(line, flush=True)

  def start(self, key, **md):
    self.event(key + '_START', None, md)

  def end(self, key, **md):
    self.event(key + '_STOP', None, md)
