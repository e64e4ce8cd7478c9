"""Hot-first feature reordering (parity: reference python/data/reorder.py:19-36)."""
import torch


def _count_ids(ids: torch.Tensor, n: int) -> torch.Tensor:
  """bincount(ids, minlength=n); large inputs are counted in slices on a few threads (the op releases the GIL) and
  summed: 20 M ids in ~45 ms instead of ~210 ms on 8 cores."""
  if ids.numel() < (1 << 21):
    return torch.bincount(ids, minlength=n)
  import os
  from concurrent.futures import ThreadPoolExecutor
  k = max(2, min(8, (os.cpu_count() or 2), torch.get_num_threads() * 2))
  with ThreadPoolExecutor(k) as pool:
    outs = list(pool.map(lambda c: torch.bincount(c, minlength=n), ids.chunk(k)))
  size = max(o.numel() for o in outs)
  out = torch.zeros(size, dtype=outs[0].dtype)
  for o in outs:
    out[:o.numel()] += o
  return out


def sort_by_in_degree(cpu_tensor: torch.Tensor, shuffle_ratio: float, topo):
  """Permute rows by descending in-degree (hot rows first); the hottest
  `shuffle_ratio` prefix is shuffled so that sharding it over a DeviceGroup is
  load-balanced.  Returns (reordered tensor, old-id -> new-row map)."""
  n = cpu_tensor.shape[0]
  if topo is None:
    return cpu_tensor, None
  if topo.layout == 'CSC':
    deg = topo.degrees  # indptr is over columns: in-degree directly
  else:
    deg = _count_ids(topo.indices, n)
  if deg.numel() < n:
    deg = torch.cat([deg, torch.zeros(n - deg.numel(), dtype=deg.dtype)])
  deg = deg[:n]
  # descending degree, ties in ascending id order == stable ASCENDING sort of (max - deg); on a narrow integer type
  # torch takes its radix path: 5 ms per million nodes instead of 65 ms for the stable descending int64 sort
  mx = int(deg.max()) if deg.numel() > 0 else 0
  narrow = torch.int16 if mx < (1 << 15) else (torch.int32 if mx < (1 << 31) else torch.int64)
  order = torch.sort((mx - deg).to(narrow), stable=True).indices
  hot = int(n * max(0.0, min(1.0, float(shuffle_ratio))))
  if hot > 1:
    order[:hot] = order[:hot][torch.randperm(hot)]
  old2new = torch.empty(n, dtype=torch.int64)
  old2new[order] = torch.arange(n, dtype=torch.int64)
  return cpu_tensor.index_select(0, order), old2new
