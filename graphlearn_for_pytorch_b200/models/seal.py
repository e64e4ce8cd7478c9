"""SEAL link prediction pieces: DRNL node labelling and a DGCNN classifier.

The reference example (examples/seal_link_pred.py:44-136) extracts an enclosing subgraph per
link with `NeighborSampler.subgraph`, labels nodes with Double-Radius Node Labelling on the
CPU (scipy shortest paths) and classifies with PyG's DGCNN.  Here DRNL is a BFS in pure
torch (runs on the sampling device) and DGCNN is a small dependency-free module.
"""
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


def _bfs_dist(edge_index: torch.Tensor, num_nodes: int, source: int, removed: Optional[int] = None,
              max_dist: int = 64) -> torch.Tensor:
  """Unweighted shortest-path distances from `source`, optionally with node `removed` deleted."""
  dev = edge_index.device
  dist = torch.full((num_nodes,), max_dist, dtype=torch.long, device=dev)
  src, dst = edge_index
  if removed is not None:
    keep = (src != removed) & (dst != removed)
    src, dst = src[keep], dst[keep]
  frontier = torch.zeros(num_nodes, dtype=torch.bool, device=dev)
  frontier[source] = True
  dist[source] = 0
  d = 0
  while bool(frontier.any()) and d < max_dist:
    d += 1
    nxt = torch.zeros(num_nodes, dtype=torch.bool, device=dev)
    nxt[dst[frontier[src]]] = True
    nxt[src[frontier[dst]]] = True          # treat the subgraph as undirected
    nxt &= dist == max_dist
    dist[nxt] = d
    frontier = nxt
  return dist


def drnl_node_labeling(edge_index: torch.Tensor, src: int, dst: int, num_nodes: int, max_z: int = 1000):
  """Double-Radius Node Labelling (Zhang & Chen, 2018): z = 1 + min(ds,dd) + (d//2)(d//2 + d%2 - 1)."""
  ds = _bfs_dist(edge_index, num_nodes, src, removed=dst)
  dd = _bfs_dist(edge_index, num_nodes, dst, removed=src)
  d = ds + dd
  half, rem = d // 2, d % 2
  z = 1 + torch.minimum(ds, dd) + half * (half + rem - 1)
  unreachable = (ds >= 64) | (dd >= 64)
  z[unreachable] = 0
  z[src] = 1
  z[dst] = 1
  return z.clamp(max=max_z)


class DGCNN(nn.Module):
  """GCN stack -> sort pooling (top-k by last channel) -> 1-D convs -> MLP."""

  def __init__(self, num_labels: int, hidden: int = 32, num_layers: int = 3, k: int = 30, emb_dim: int = 32):
    super().__init__()
    self.k = k
    self.z_emb = nn.Embedding(num_labels, emb_dim)
    dims = [emb_dim] + [hidden] * num_layers + [1]
    self.lins = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)])
    total = hidden * num_layers + 1
    self.conv1 = nn.Conv1d(1, 16, total, total)
    self.pool = nn.MaxPool1d(2, 2)
    self.conv2 = nn.Conv1d(16, 32, 5, 1)
    dense = int((k - 2) / 2 + 1)
    self.dense_dim = max((dense - 5 + 1) * 32, 32)
    self.lin1 = nn.Linear(self.dense_dim, 128)
    self.lin2 = nn.Linear(128, 1)

  def _gcn(self, x, edge_index, lin):
    n = x.shape[0]
    loops = torch.arange(n, device=x.device)
    src = torch.cat([edge_index[0], loops])
    dst = torch.cat([edge_index[1], loops])
    deg = torch.zeros(n, device=x.device).index_add_(0, dst, torch.ones_like(dst, dtype=torch.float))
    norm = (deg[src] * deg[dst]).clamp(min=1).rsqrt()
    h = lin(x)
    return torch.zeros_like(h).index_add_(0, dst, h[src] * norm.unsqueeze(1))

  def forward(self, z: torch.Tensor, edge_index: torch.Tensor, batch: torch.Tensor, num_graphs: int):
    x = self.z_emb(z)
    xs = []
    for lin in self.lins:
      x = torch.tanh(self._gcn(x, edge_index, lin))
      xs.append(x)
    x = torch.cat(xs, dim=-1)
    # sort pooling per graph
    out = x.new_zeros(num_graphs, self.k, x.shape[1])
    for g in range(num_graphs):
      xg = x[batch == g]
      order = torch.argsort(xg[:, -1], descending=True)[:self.k]
      out[g, :order.numel()] = xg[order]
    h = out.view(num_graphs, 1, -1)
    h = F.relu(self.conv1(h))
    h = self.pool(h)
    if h.shape[-1] >= 5:
      h = F.relu(self.conv2(h))
    h = h.reshape(num_graphs, -1)
    if h.shape[1] != self.dense_dim:
      h = F.pad(h, (0, max(0, self.dense_dim - h.shape[1])))[:, :self.dense_dim]
    h = F.dropout(F.relu(self.lin1(h)), 0.5, self.training)
    return self.lin2(h).view(-1)
