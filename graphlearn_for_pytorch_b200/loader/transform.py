"""SamplerOutput -> Data / HeteroData (parity: reference python/loader/transform.py:26-136)."""
from typing import Dict, Literal, Optional

import torch

from ..sampler import HeteroSamplerOutput, SamplerOutput
from ..typing import EdgeType, NodeType, reverse_edge_type
from .data import Data, HeteroData


def to_data(sampler_out: SamplerOutput, batch_labels: Optional[torch.Tensor] = None,
            node_feats: Optional[torch.Tensor] = None, edge_feats: Optional[torch.Tensor] = None,
            **kwargs) -> Data:
  edge_index = torch.stack([sampler_out.row, sampler_out.col])
  data = Data(x=node_feats, edge_index=edge_index, edge_attr=edge_feats, y=batch_labels, **kwargs)
  data.edge = sampler_out.edge
  data.node = sampler_out.node
  data.batch = sampler_out.batch
  data.batch_size = sampler_out.batch.numel() if sampler_out.batch is not None else 0
  data.num_sampled_nodes = sampler_out.num_sampled_nodes
  data.num_sampled_edges = sampler_out.num_sampled_edges
  md = sampler_out.metadata
  if isinstance(md, dict):
    for k, v in md.items():
      if k == 'edge_label_index':
        # sampled edges point neighbour -> seed, so the supervision pairs are flipped too
        data['edge_label_index'] = torch.stack((v[1], v[0]), dim=0)
      else:
        data[k] = v
  elif md is not None:
    data['metadata'] = md
    data['mapping'] = md
  return data


def _pad(counts, n, device):
  t = torch.as_tensor(counts, dtype=torch.int64)
  if t.numel() < n:
    t = torch.cat([t, torch.zeros(n - t.numel(), dtype=torch.int64)])
  return t


def to_hetero_data(hetero_sampler_out: HeteroSamplerOutput,
                   batch_label_dict: Optional[Dict[NodeType, torch.Tensor]] = None,
                   node_feat_dict: Optional[Dict[NodeType, torch.Tensor]] = None,
                   edge_feat_dict: Optional[Dict[EdgeType, torch.Tensor]] = None,
                   edge_dir: Literal['in', 'out'] = 'out', **kwargs) -> HeteroData:
  out = hetero_sampler_out
  data = HeteroData(**kwargs)
  edge_index_dict = out.get_edge_index()
  nse = out.num_sampled_edges or {}
  nsn = out.num_sampled_nodes or {}
  num_hops = max([len(v) for v in nse.values()] + [0])
  for k, v in edge_index_dict.items():
    data[k].edge_index = v
    if out.edge is not None:
      data[k].edge = out.edge.get(k)
    if edge_feat_dict is not None:
      data[k].edge_attr = edge_feat_dict.get(k)
    nse[k] = _pad(nse.get(k, []), num_hops, v.device)
  for k, v in out.node.items():
    data[k].node = v
    if node_feat_dict is not None:
      data[k].x = node_feat_dict.get(k)
    nsn[k] = _pad(nsn.get(k, []), num_hops + 1, v.device)
  for k, v in (out.batch or {}).items():
    data[k].batch = v
    data[k].batch_size = v.numel()
  if batch_label_dict is not None:
    for k, v in batch_label_dict.items():
      data[k].y = v
  data.num_sampled_nodes = nsn
  data.num_sampled_edges = nse
  input_type = out.input_type
  md = out.metadata
  if isinstance(md, dict):
    res_type = reverse_edge_type(input_type) if edge_dir == 'out' else input_type
    for k, v in md.items():
      if k == 'edge_label_index':
        data[res_type].edge_label_index = torch.stack((v[1], v[0]), dim=0) if edge_dir == 'out' else v
      elif k == 'edge_label':
        data[res_type].edge_label = v
      elif k == 'src_index':
        data[input_type[0]].src_index = v
      elif k in ('dst_pos_index', 'dst_neg_index'):
        data[input_type[-1]][k] = v
      else:
        data[k] = v
  elif md is not None:
    data['metadata'] = md
  return data
