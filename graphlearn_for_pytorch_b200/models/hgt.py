"""Heterogeneous Graph Transformer (Hu et al., 2020) for `HeteroData` mini-batches.

The reference trains PyG's HGTConv on OGB-MAG (examples/hetero/train_hgt_mag.py); this is a
dependency-free implementation of the same layer: per-node-type K/Q/V projections,
per-relation attention (W_att) and message (W_msg) transforms with a learnable relation prior,
softmax over *all* incoming edges of a target node (across relations), gated skip connection.
"""
import math
from typing import Dict, List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

EdgeType = Tuple[str, str, str]


class HGTConv(nn.Module):
  """Heterogeneous Graph Transformer layer: per-type K/Q/V projections, per-relation attention and message
  transforms with relation priors, softmax over the incoming edges of every target, gated skip connection."""
  def __init__(self, in_channels: Dict[str, int], out_channels: int, node_types: List[str],
               edge_types: List[EdgeType], heads: int = 4):
    super().__init__()
    assert out_channels % heads == 0
    self.h, self.d = heads, out_channels // heads
    self.node_types, self.edge_types = list(node_types), [tuple(e) for e in edge_types]
    self.k = nn.ModuleDict({t: nn.Linear(in_channels[t], out_channels) for t in node_types})
    self.q = nn.ModuleDict({t: nn.Linear(in_channels[t], out_channels) for t in node_types})
    self.v = nn.ModuleDict({t: nn.Linear(in_channels[t], out_channels) for t in node_types})
    self.a = nn.ModuleDict({t: nn.Linear(out_channels, out_channels) for t in node_types})
    self.skip = nn.ParameterDict({t: nn.Parameter(torch.ones(1)) for t in node_types})
    self.res = nn.ModuleDict({t: (nn.Identity() if in_channels[t] == out_channels else
                                  nn.Linear(in_channels[t], out_channels, bias=False)) for t in node_types})
    key = lambda e: '__'.join(e)
    self.w_att = nn.ParameterDict({key(e): nn.Parameter(torch.randn(heads, self.d, self.d) / math.sqrt(self.d))
                                   for e in self.edge_types})
    self.w_msg = nn.ParameterDict({key(e): nn.Parameter(torch.randn(heads, self.d, self.d) / math.sqrt(self.d))
                                   for e in self.edge_types})
    self.prior = nn.ParameterDict({key(e): nn.Parameter(torch.ones(heads)) for e in self.edge_types})

  def forward(self, x_dict: Dict[str, torch.Tensor], edge_index_dict: Dict[EdgeType, torch.Tensor]):
    H, D = self.h, self.d
    k = {t: self.k[t](x).view(-1, H, D) for t, x in x_dict.items() if t in self.k}
    q = {t: self.q[t](x).view(-1, H, D) for t, x in x_dict.items() if t in self.q}
    v = {t: self.v[t](x).view(-1, H, D) for t, x in x_dict.items() if t in self.v}
    # gather every relation's edges per destination type so that the softmax spans all of them
    per_dst: Dict[str, list] = {}
    for et in self.edge_types:
      ei = edge_index_dict.get(et)
      if ei is None or ei.numel() == 0 or et[0] not in k or et[2] not in q:
        continue
      name = '__'.join(et)
      src, dst = ei[0], ei[1]
      kk = torch.einsum('ehd,hdf->ehf', k[et[0]][src], self.w_att[name])
      score = (q[et[2]][dst] * kk).sum(-1) * self.prior[name] / math.sqrt(D)          # [E, H]
      msg = torch.einsum('ehd,hdf->ehf', v[et[0]][src], self.w_msg[name])             # [E, H, D]
      per_dst.setdefault(et[2], []).append((dst, score, msg))
    out = {}
    for t, x in x_dict.items():
      if t not in self.a:
        continue
      n = x.shape[0]
      if t in per_dst:
        dst = torch.cat([p[0] for p in per_dst[t]])
        score = torch.cat([p[1] for p in per_dst[t]]).float()     # fp32 softmax / weighted sum also under autocast
        msg = torch.cat([p[2] for p in per_dst[t]])
        mx = torch.full((n, H), -1e30, dtype=score.dtype, device=score.device)
        mx = mx.scatter_reduce(0, dst.unsqueeze(1).expand(-1, H), score, reduce='amax')
        e = torch.exp(score - mx[dst])
        den = torch.zeros(n, H, dtype=e.dtype, device=e.device).index_add_(0, dst, e)
        w = e / den[dst].clamp(min=1e-16)
        agg = torch.zeros(n, H, D, dtype=torch.float32, device=msg.device).index_add_(0, dst,
                                                                                      msg.float() * w.unsqueeze(-1))
        h = self.a[t](F.gelu(agg.reshape(n, H * D).to(msg.dtype)))
      else:
        h = torch.zeros(n, H * D, dtype=x.dtype, device=x.device)
      alpha = torch.sigmoid(self.skip[t])
      out[t] = alpha * h + (1 - alpha) * self.res[t](x)
    return out


class HGT(nn.Module):
  """Stack of HGTConv layers + a linear head on `node_type`."""

  def __init__(self, node_types: List[str], edge_types: List[EdgeType], in_channels, hidden_channels: int,
               out_channels: int, num_layers: int = 2, heads: int = 4, node_type: str = 'paper'):
    super().__init__()
    in_dim = in_channels if isinstance(in_channels, dict) else {t: in_channels for t in node_types}
    self.node_type = node_type
    self.convs = nn.ModuleList()
    for l in range(num_layers):
      dims = in_dim if l == 0 else {t: hidden_channels for t in node_types}
      self.convs.append(HGTConv(dims, hidden_channels, node_types, edge_types, heads))
    self.head = nn.Linear(hidden_channels, out_channels)

  def forward(self, x_dict, edge_index_dict):
    for conv in self.convs:
      x_dict = conv(x_dict, edge_index_dict)
    return self.head(x_dict[self.node_type])
