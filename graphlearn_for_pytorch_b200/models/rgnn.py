"""Heterogeneous relational GNNs (R-GCN / R-SAGE / R-GAT) for `HeteroData` mini-batches.

The reference keeps its hetero model in the IGBH example (examples/igbh/rgnn.py:22-81: a PyG
HeteroConv of SAGEConv / GATConv per edge type with trim_to_layer).  This module is the
dependency-free equivalent: one relation-specific convolution per edge type, summed per
destination node type, optional per-layer trimming driven by num_sampled_nodes/edges.
"""
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

EdgeType = Tuple[str, str, str]


def _segment_mean(src_feat: torch.Tensor, dst_index: torch.Tensor, n_dst: int) -> torch.Tensor:
  out = torch.zeros(n_dst, src_feat.shape[1], dtype=src_feat.dtype, device=src_feat.device)
  out.index_add_(0, dst_index, src_feat)
  # degrees are counted in fp32: bf16 holds integers exactly only up to 256
  if dst_index.is_cuda:   # bincount synchronises with the host on CUDA (it needs max()); index_add does not
    deg = torch.zeros(n_dst, dtype=torch.float32, device=src_feat.device)
    deg.index_add_(0, dst_index, torch.ones_like(dst_index, dtype=torch.float32))
  else:
    deg = torch.bincount(dst_index, minlength=n_dst).to(torch.float32)
  return out * (1.0 / deg.clamp(min=1)).to(out.dtype).unsqueeze(1)


class RelSAGEConv(nn.Module):
  """mean_{j in N_r(i)} x_j W_r  +  x_i W_self   for one relation r: src_type -> dst_type."""

  def __init__(self, in_src: int, in_dst: int, out: int):
    super().__init__()
    self.lin_l = nn.Linear(in_src, out)
    self.lin_r = nn.Linear(in_dst, out, bias=False)

  def forward(self, x_src, x_dst, edge_index):
    if self.lin_l.in_features > self.lin_l.out_features:
      # the mean is linear: project the (fewer, wider) source rows first, so that the per-edge gather moves
      # `out`-wide rows instead of `in`-wide ones (IGBH: 1024 -> 256, 4x less gather traffic); the bias is
      # added after the aggregation so that isolated targets still get exactly lin_l(0) = b
      proj = F.linear(x_src, self.lin_l.weight)
      agg = _segment_mean(proj[edge_index[0]], edge_index[1], x_dst.shape[0])
      if self.lin_l.bias is not None:
        agg = agg + self.lin_l.bias
      return agg + self.lin_r(x_dst)
    agg = _segment_mean(x_src[edge_index[0]], edge_index[1], x_dst.shape[0])
    return self.lin_l(agg) + self.lin_r(x_dst)


class RelGCNConv(nn.Module):
  """R-GCN relation: (1/c_i) sum_j x_j W_r  (self loop handled by the caller's root weight)."""

  def __init__(self, in_src: int, in_dst: int, out: int):
    super().__init__()
    self.lin = nn.Linear(in_src, out, bias=False)

  def forward(self, x_src, x_dst, edge_index):
    return _segment_mean(self.lin(x_src)[edge_index[0]], edge_index[1], x_dst.shape[0])


class RelGATConv(nn.Module):
  """Single-relation multi-head attention (GATConv semantics, softmax over incoming edges)."""

  def __init__(self, in_src: int, in_dst: int, out: int, heads: int = 4):
    super().__init__()
    assert out % heads == 0
    self.h, self.c = heads, out // heads
    self.lin_src = nn.Linear(in_src, out, bias=False)
    self.lin_dst = nn.Linear(in_dst, out, bias=False)
    self.att_src = nn.Parameter(torch.randn(heads, self.c) * 0.1)
    self.att_dst = nn.Parameter(torch.randn(heads, self.c) * 0.1)
    self.bias = nn.Parameter(torch.zeros(out))

  def forward(self, x_src, x_dst, edge_index):
    n_dst = x_dst.shape[0]
    hs = self.lin_src(x_src).view(-1, self.h, self.c)
    hd = self.lin_dst(x_dst).view(-1, self.h, self.c)
    a = (hs * self.att_src).sum(-1)[edge_index[0]] + (hd * self.att_dst).sum(-1)[edge_index[1]]
    # scores, softmax and the weighted sum run in fp32 whatever the activations' dtype is (bf16 under autocast):
    # the index_add_ operands then agree and the normalisation keeps its precision
    a = F.leaky_relu(a, 0.2).float()
    amax = torch.full((n_dst, self.h), -1e30, dtype=a.dtype, device=a.device)
    amax = amax.scatter_reduce(0, edge_index[1].unsqueeze(1).expand(-1, self.h), a, reduce='amax')
    e = torch.exp(a - amax[edge_index[1]])
    denom = torch.zeros(n_dst, self.h, dtype=a.dtype, device=a.device).index_add_(0, edge_index[1], e)
    w = e / denom[edge_index[1]].clamp(min=1e-16)
    out = torch.zeros(n_dst, self.h, self.c, dtype=torch.float32, device=hs.device)
    out.index_add_(0, edge_index[1], hs[edge_index[0]].float() * w.unsqueeze(-1))
    return out.reshape(n_dst, -1).to(hs.dtype) + self.bias


_CONVS = {'rsage': RelSAGEConv, 'rgcn': RelGCNConv, 'rgat': RelGATConv}


class RGNN(nn.Module):
  """Args:
    edge_types: relations *as they appear in the sampled batch* (message direction src->dst).
    in_channels: input width (int, or dict per node type).
    hidden_channels / out_channels / num_layers: model size.
    node_type: the type whose logits are returned (seed type).
    model: 'rsage' | 'rgcn' | 'rgat'.
  """

  def __init__(self, edge_types: List[EdgeType], in_channels, hidden_channels: int, out_channels: int,
               num_layers: int = 2, node_type: str = 'paper', model: str = 'rsage', dropout: float = 0.2,
               heads: int = 4):
    super().__init__()
    self.edge_types = [tuple(e) for e in edge_types]
    self.node_type = node_type
    self.dropout = dropout
    ntypes = sorted({t for e in self.edge_types for t in (e[0], e[2])})
    in_dim = in_channels if isinstance(in_channels, dict) else {t: in_channels for t in ntypes}
    self.layers = nn.ModuleList()
    self.roots = nn.ModuleList()
    conv_cls = _CONVS[model]
    for l in range(num_layers):
      dims = in_dim if l == 0 else {t: hidden_channels for t in ntypes}
      out = hidden_channels if l < num_layers - 1 else out_channels
      convs = nn.ModuleDict()
      for (s, r, d) in self.edge_types:
        kw = {'heads': heads} if model == 'rgat' and out % heads == 0 else {}
        cls = conv_cls if not (model == 'rgat' and out % heads != 0) else RelSAGEConv
        convs['__'.join((s, r, d))] = cls(dims[s], dims[d], out, **kw)
      self.layers.append(convs)
      # R-GCN needs an explicit self/root transform per destination type
      self.roots.append(nn.ModuleDict({t: nn.Linear(dims[t], out) for t in ntypes}) if model == 'rgcn' else
                        nn.ModuleDict())

  def forward(self, x_dict: Dict[str, torch.Tensor], edge_index_dict: Dict[EdgeType, torch.Tensor],
              num_sampled_nodes_dict: Optional[Dict[str, List[int]]] = None,
              num_sampled_edges_dict: Optional[Dict[EdgeType, List[int]]] = None):
    L = len(self.layers)
    for l, convs in enumerate(self.layers):
      if num_sampled_nodes_dict is not None and num_sampled_edges_dict is not None:
        keep = L - l   # hops still needed (trim_to_layer)
        n_keep = {t: int(sum(list(v)[:keep + 1])) for t, v in num_sampled_nodes_dict.items()}
        x_dict = {t: x[:n_keep.get(t, x.shape[0])] for t, x in x_dict.items()}
        edge_index_dict = {et: ei[:, :int(sum(list(num_sampled_edges_dict.get(et, [ei.shape[1]]))[:keep]))]
                           for et, ei in edge_index_dict.items()}
      out: Dict[str, torch.Tensor] = {}
      for et in self.edge_types:
        ei = edge_index_dict.get(et)
        if ei is None or et[0] not in x_dict or et[2] not in x_dict:
          continue
        h = convs['__'.join(et)](x_dict[et[0]], x_dict[et[2]], ei)
        out[et[2]] = h if et[2] not in out else out[et[2]] + h
      for t, root in self.roots[l].items():
        if t in x_dict:
          out[t] = root(x_dict[t]) + (out[t] if t in out else 0)
      if l < L - 1:
        out = {t: F.dropout(F.leaky_relu(h), self.dropout, self.training) for t, h in out.items()}
      # node types without incoming relations in this layer keep a zero state of the right width
      width = next(iter(out.values())).shape[1] if out else 0
      for t, x in x_dict.items():
        if t not in out and width:
          out[t] = torch.zeros(x.shape[0], width, dtype=x.dtype, device=x.device)
      x_dict = out
    return x_dict[self.node_type]
