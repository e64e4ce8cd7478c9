"""Shared fixtures: tiny self-describing graphs (feature row v == [v]*dim)."""
import torch

import graphlearn_for_pytorch_b200 as glt
from graphlearn_for_pytorch_b200.utils.synthetic import id_features, ring_graph, rmat_edges


def ring_dataset(num_nodes=40, dim=16, graph_mode='CPU', edge_dir='out', with_gpu=False, device=None,
                 split_ratio=0.0, weights=False):
  ei = ring_graph(num_nodes)
  ds = glt.data.Dataset(edge_dir=edge_dir)
  w = (torch.arange(ei.shape[1], dtype=torch.float32) % 3 + 1.0) if weights else None
  ds.init_graph(ei, edge_weights=w, graph_mode=graph_mode, device=device, num_nodes=num_nodes)
  ds.init_node_features(id_features(num_nodes, dim), with_gpu=with_gpu, split_ratio=split_ratio, device=device)
  ds.init_edge_features(id_features(ei.shape[1], 4), with_gpu=with_gpu, device=device)
  ds.init_node_labels(torch.arange(num_nodes))
  return ds


def rmat_csr(num_nodes=2000, num_edges=40000, seed=0):
  ei = rmat_edges(num_nodes, num_edges, seed=seed)
  topo = glt.data.Topology(ei, layout='CSR', num_nodes=num_nodes)
  return ei, topo


def adjacency_sets(topo):
  ptr, ind = topo.indptr.tolist(), topo.indices.tolist()
  return [set(ind[ptr[i]:ptr[i + 1]]) for i in range(len(ptr) - 1)]


def canonical_edges(out):
  """Set of (global src, global dst) pairs of a SamplerOutput."""
  node = out.node.cpu()
  return set(zip(node[out.row.cpu()].tolist(), node[out.col.cpu()].tolist()))
