import torch

import graphlearn_for_pytorch_b200 as glt
from graphlearn_for_pytorch_b200.loader import (LinkNeighborLoader, NeighborLoader, SubGraphLoader)
from graphlearn_for_pytorch_b200.sampler import NegativeSampling
from helpers import ring_dataset


def check_homo_batch(b, n=40):
  assert torch.equal(b.x[:, 0].long(), b.node)            # feature row == node id
  assert torch.equal(b.y, b.node)                          # label == node id
  src, dst = b.node[b.edge_index[0]], b.node[b.edge_index[1]]
  assert torch.all(((src - dst) % n == 1) | ((src - dst) % n == 2))
  assert sum(b.num_sampled_nodes) == b.node.numel()
  assert sum(b.num_sampled_edges) == b.edge_index.shape[1]
  assert b.batch_size == b.batch.numel()
  if b.edge is not None and b.edge_attr is not None:
    assert torch.equal(b.edge_attr[:, 0].long(), b.edge)


def test_neighbor_loader_epochs():
  ds = ring_dataset(40)
  loader = NeighborLoader(ds, [2, 2], torch.arange(40), batch_size=8, shuffle=True, drop_last=False,
                          with_edge=True, seed=7)
  assert len(loader) == 5
  for _ in range(2):
    seen = []
    for b in loader:
      check_homo_batch(b)
      seen += b.batch.tolist()
    assert sorted(seen) == list(range(40))


def test_neighbor_loader_checkpoint_resume():
  ds = ring_dataset(40)
  mk = lambda: NeighborLoader(ds, [1, 1], torch.arange(40), batch_size=8, shuffle=True, seed=3)
  a = mk()
  it = iter(a)
  next(it); next(it)
  state = a.state_dict()
  rest_a = []
  while True:
    try:
      rest_a.append(next(a).node.tolist())
    except StopIteration:
      break
  b = mk()
  iter(b)
  b.load_state_dict(state)
  rest_b = []
  while True:
    try:
      rest_b.append(next(b).node.tolist())
    except StopIteration:
      break
  assert rest_a == rest_b


def test_pyg_v1_loader():
  ds = ring_dataset(40)
  loader = NeighborLoader(ds, [2, 2], torch.arange(16), batch_size=8, as_pyg_v1=True)
  for bs, n_id, adjs in loader:
    assert bs == 8 and len(adjs) == 2
    x = ds.node_features[n_id]
    assert torch.equal(x[:, 0].long(), n_id)


def test_link_neighbor_loader_binary_and_triplet():
  ds = ring_dataset(40)
  ei = torch.stack(ds.graph.topo.to_coo()[:2])
  loader = LinkNeighborLoader(ds, [2], edge_label_index=ei, neg_sampling=NegativeSampling('binary', 1),
                              batch_size=10, shuffle=True, seed=1)
  n = 0
  for b in loader:
    eli, lab = b.edge_label_index, b.edge_label
    assert eli.shape[1] == lab.numel()
    pos = lab == 1
    # to_data flips edge_label_index: row 0 = dst, row 1 = src
    src, dst = b.node[eli[1]], b.node[eli[0]]
    assert torch.all(((dst[pos] - src[pos]) % 40 == 1) | ((dst[pos] - src[pos]) % 40 == 2))
    assert torch.all(((dst[~pos] - src[~pos]) % 40 != 1) & ((dst[~pos] - src[~pos]) % 40 != 2))
    assert torch.equal(b.x[:, 0].long(), b.node)
    n += int(pos.sum())
  assert n == 80
  loader = LinkNeighborLoader(ds, [2], edge_label_index=ei, neg_sampling=NegativeSampling('triplet', 2),
                              batch_size=16)
  for b in loader:
    assert b.dst_neg_index.shape == (b.src_index.numel(), 2)
    src, dst = b.node[b.src_index], b.node[b.dst_pos_index]
    assert torch.all(((dst - src) % 40 == 1) | ((dst - src) % 40 == 2))


def test_subgraph_loader():
  ds = ring_dataset(40)
  loader = SubGraphLoader(ds, torch.arange(0, 40, 10), num_neighbors=[-1], batch_size=2, with_edge=True)
  for b in loader:
    assert torch.equal(b.node[b.mapping], b.batch)
    assert torch.equal(b.x[:, 0].long(), b.node)
    src, dst = b.node[b.edge_index[0]], b.node[b.edge_index[1]]
    assert torch.all(((src - dst) % 40 == 1) | ((src - dst) % 40 == 2))      # same orientation as NeighborLoader batches


def test_hetero_neighbor_loader():
  u2i = torch.tensor([[0, 0, 1, 2, 3], [0, 1, 1, 2, 3]])
  i2i = torch.tensor([[0, 1, 2, 3], [1, 2, 3, 0]])
  ds = glt.data.Dataset(edge_dir='out')
  ds.init_graph({('user', 'u2i', 'item'): u2i, ('item', 'i2i', 'item'): i2i}, graph_mode='CPU')
  ds.init_node_features({'user': glt.utils.id_features(4, 8), 'item': glt.utils.id_features(4, 8) + 100},
                        with_gpu=False)
  ds.init_node_labels({'user': torch.arange(4)})
  loader = NeighborLoader(ds, [2, 2], ('user', torch.arange(4)), batch_size=2)
  for b in loader:
    assert torch.equal(b['user'].x[:, 0].long(), b['user'].node)
    assert torch.equal(b['item'].x[:, 0].long() - 100, b['item'].node)
    assert b['user'].batch_size == 2 and torch.equal(b['user'].y[:2], b['user'].batch)
    ei = b['item', 'rev_u2i', 'user'].edge_index
    assert ei.shape[0] == 2 and ei[1].max() < b['user'].node.numel()
    assert set(b.edge_index_dict.keys()) <= {('item', 'rev_u2i', 'user'), ('item', 'i2i', 'item')}
    assert len(b.num_sampled_nodes['item']) == 3


def _hetero_ds(edge_dir='out'):
  u2i = torch.tensor([[0, 0, 1, 2, 3, 3, 1], [0, 1, 1, 2, 3, 0, 2]])
  i2i = torch.tensor([[0, 1, 2, 3], [1, 2, 3, 0]])
  ds = glt.data.Dataset(edge_dir=edge_dir)
  ds.init_graph({('user', 'u2i', 'item'): u2i, ('item', 'i2i', 'item'): i2i}, graph_mode='CPU',
                num_nodes={'user': 4, 'item': 4})
  ds.init_node_features({'user': glt.utils.id_features(4, 8), 'item': glt.utils.id_features(4, 8) + 100},
                        with_gpu=False)
  return ds, u2i


def test_hetero_link_neighbor_loader_binary_and_triplet():
  ds, u2i = _hetero_ds()
  et = ('user', 'u2i', 'item')
  loader = LinkNeighborLoader(ds, [2, 2], edge_label_index=(et, u2i), neg_sampling=NegativeSampling('binary', 1),
                              batch_size=4, shuffle=True, seed=0)
  pos_edges = set(zip(u2i[0].tolist(), u2i[1].tolist()))
  n_pos = 0
  for b in loader:
    rev = ('item', 'rev_u2i', 'user')
    eli, lab = b[rev].edge_label_index, b[rev].edge_label
    # reversed relation: row 0 = item (dst), row 1 = user (src)
    users, items = b['user'].node[eli[1]], b['item'].node[eli[0]]
    for u, i, l in zip(users.tolist(), items.tolist(), lab.tolist()):
      assert ((u, i) in pos_edges) == (l == 1)
    assert torch.equal(b['user'].x[:, 0].long(), b['user'].node)
    assert torch.equal(b['item'].x[:, 0].long() - 100, b['item'].node)
    n_pos += int((lab == 1).sum())
  assert n_pos == u2i.shape[1]
  loader = LinkNeighborLoader(ds, [2], edge_label_index=(et, u2i), neg_sampling=NegativeSampling('triplet', 2),
                              batch_size=4)
  for b in loader:
    src = b['user'].node[b['user'].src_index]
    dst = b['item'].node[b['item'].dst_pos_index]
    assert all((u, i) in pos_edges for u, i in zip(src.tolist(), dst.tolist()))
    assert b['item'].dst_neg_index.shape == (src.numel(), 2)


def test_table_dataset_from_columns():
  import numpy as np
  from graphlearn_for_pytorch_b200.data import TableDataset
  edges = {'src_id': np.array([0, 1, 2, 3]), 'dst_id': np.array([1, 2, 3, 0])}
  nodes = {'id': np.arange(4), 'feature': np.array(['0:0', '1:1', '2:2', '3:3'], dtype=object),
           'label': np.array([0, 1, 0, 1])}
  ds = TableDataset().load(edge_tables={('n', 'e', 'n'): edges}, node_tables={'n': nodes}, graph_mode='CPU',
                           directed=True)
  loader = NeighborLoader(ds, [1], torch.arange(4), batch_size=2)
  for b in loader:
    assert torch.equal(b.x[:, 0].long(), b.node) and torch.equal(b.y, b.node % 2)
    src, dst = b.node[b.edge_index[0]], b.node[b.edge_index[1]]
    assert torch.all((src - dst) % 4 == 1)
