import torch

import graphlearn_for_pytorch_b200 as glt
from graphlearn_for_pytorch_b200.sampler import (EdgeSamplerInput, NegativeSampling, NeighborSampler,
                                                  NodeSamplerInput)
from helpers import canonical_edges, ring_dataset


def test_two_hop_ring_exact():
  """BASELINE config 1: 2-hop sampling on a tiny CSR, CPU, world_size=1."""
  ds = ring_dataset(40)
  s = NeighborSampler(ds.graph, [2, 2], with_edge=True, seed=1)
  out = s.sample_from_nodes(NodeSamplerInput(torch.tensor([0, 10])))
  # deg == fanout: sampling is deterministic
  assert out.node[:2].tolist() == [0, 10]
  assert set(out.node.tolist()) == {0, 1, 2, 3, 4, 10, 11, 12, 13, 14}
  assert out.num_sampled_nodes == [2, 4, 4]
  assert out.num_sampled_edges == [4, 8]
  edges = canonical_edges(out)          # (neighbour, seed) pairs
  assert edges == {(1, 0), (2, 0), (11, 10), (12, 10), (2, 1), (3, 1), (3, 2), (4, 2),
                   (12, 11), (13, 11), (13, 12), (14, 12)}
  src, dst = out.node[out.row], out.node[out.col]
  assert torch.all(((src - dst) % 40 == 1) | ((src - dst) % 40 == 2))
  # edge ids are the ids of (dst -> src) edges in the ring
  r, c, e, _ = ds.graph.topo.to_coo()
  eid_of = {(a, b): i for a, b, i in zip(r.tolist(), c.tolist(), e.tolist())}
  for a, b, i in zip(dst.tolist(), src.tolist(), out.edge.tolist()):
    assert eid_of[(a, b)] == i


def test_duplicate_seeds_and_early_stop():
  ds = ring_dataset(40)
  s = NeighborSampler(ds.graph, [2], seed=1)
  out = s.sample_from_nodes(torch.tensor([5, 5, 6]))
  assert out.node[:2].tolist() == [5, 6] and out.batch.tolist() == [5, 6]
  # isolated node: no edges -> hops stop
  topo = glt.data.Topology(torch.tensor([[0], [1]]), layout='CSR', num_nodes=3)
  g = glt.data.Graph(topo, 'CPU')
  out = NeighborSampler(g, [2, 2]).sample_from_nodes(torch.tensor([2]))
  assert out.node.tolist() == [2] and out.row.numel() == 0


def test_one_hop_api_and_all_neighbors():
  ds = ring_dataset(40)
  s = NeighborSampler(ds.graph, [-1], with_edge=True)
  o = s.sample_one_hop(torch.tensor([3, 39]), -1)
  assert o.nbr.tolist() == [4, 5, 0, 1] and o.nbr_num.tolist() == [2, 2] and o.edge.numel() == 4
  o = s.sample_one_hop(torch.tensor([3]), 1)
  assert o.nbr_num.tolist() == [1] and o.nbr.item() in (4, 5)


def test_in_direction():
  ds = ring_dataset(40, edge_dir='in')
  s = NeighborSampler(ds.graph, [2], edge_dir='in')
  out = s.sample_from_nodes(torch.tensor([10]))
  src, dst = out.node[out.row], out.node[out.col]
  # in-neighbours of 10 are 8 and 9; edges keep their true direction src -> dst
  assert set(src.tolist()) == {8, 9} and set(dst.tolist()) == {10}


def test_pyg_v1_triple():
  ds = ring_dataset(40)
  s = NeighborSampler(ds.graph, [2, 2])
  bs, n_id, adjs = s.sample_pyg_v1(torch.tensor([0, 20]))
  assert bs == 2 and len(adjs) == 2
  assert adjs[0].size == (n_id.numel(), 6) and adjs[1].size == (6, 2)
  assert adjs[1].edge_index[1].max() < 2 and adjs[0].edge_index[1].max() < 6


def test_link_sampling_binary_and_triplet():
  ds = ring_dataset(40)
  s = NeighborSampler(ds.graph, [2], with_neg=True, seed=3)
  row, col = torch.tensor([0, 1, 2]), torch.tensor([1, 3, 3])
  out = s.sample_from_edges(EdgeSamplerInput(row, col, neg_sampling=NegativeSampling('binary', 2)))
  eli, lab = out.metadata['edge_label_index'], out.metadata['edge_label']
  assert eli.shape == (2, 9) and lab.tolist() == [1, 1, 1] + [0] * 6
  assert out.node[eli[0, :3]].tolist() == [0, 1, 2] and out.node[eli[1, :3]].tolist() == [1, 3, 3]
  neg_src, neg_dst = out.node[eli[0, 3:]], out.node[eli[1, 3:]]
  assert torch.all(((neg_dst - neg_src) % 40 != 1) & ((neg_dst - neg_src) % 40 != 2))
  out = s.sample_from_edges(EdgeSamplerInput(row, col, neg_sampling=NegativeSampling('triplet', 2)))
  md = out.metadata
  assert out.node[md['src_index']].tolist() == [0, 1, 2]
  assert out.node[md['dst_pos_index']].tolist() == [1, 3, 3]
  assert md['dst_neg_index'].shape == (3, 2)


def test_subgraph_and_mapping():
  ds = ring_dataset(40)
  s = NeighborSampler(ds.graph, [-1], with_edge=True)
  out = s.subgraph(NodeSamplerInput(torch.tensor([0, 3])))
  nodes = set(out.node.tolist())
  assert nodes == {0, 1, 2, 3, 4, 5}
  assert out.node[out.metadata].tolist() == [0, 3]
  assert out.node.tolist() == sorted(nodes)                     # ascending ids (reference convention)
  # reversed edge index, like sample_from_nodes: row = the adjacency's column side (a + 1, a + 2), col = a
  nbr, src = out.node[out.row], out.node[out.col]
  got = set(zip(src.tolist(), nbr.tolist()))
  want = {(a, b) for a in nodes for b in nodes if (b - a) % 40 in (1, 2)}
  assert got == want


def test_sample_prob_monotone():
  ds = ring_dataset(40)
  s = NeighborSampler(ds.graph, [2, 2])
  p = s.sample_prob(NodeSamplerInput(torch.tensor([0])), 40)
  assert p.shape == (40,) and float(p.max()) <= 1.0 + 1e-6


def test_state_dict_replay():
  ds = ring_dataset(40)
  s = NeighborSampler(ds.graph, [1, 1], seed=5)
  st = s.state_dict()
  a = s.sample_from_nodes(torch.arange(10))
  s.load_state_dict(st)
  b = s.sample_from_nodes(torch.arange(10))
  assert torch.equal(a.node, b.node) and torch.equal(a.row, b.row)


def test_hetero_sampler():
  # user -u2i-> item, item -i2i-> item
  u2i = torch.tensor([[0, 0, 1, 2], [0, 1, 1, 2]])
  i2i = torch.tensor([[0, 1, 2], [1, 2, 0]])
  ds = glt.data.Dataset(edge_dir='out')
  ds.init_graph({('user', 'u2i', 'item'): u2i, ('item', 'i2i', 'item'): i2i}, graph_mode='CPU')
  s = NeighborSampler(ds.graph, [2, 2], with_edge=True)
  out = s.sample_from_nodes(NodeSamplerInput(torch.tensor([0]), 'user'))
  assert out.node['user'].tolist() == [0]
  assert set(out.node['item'].tolist()) == {0, 1, 2}
  k1, k2 = ('item', 'rev_u2i', 'user'), ('item', 'i2i', 'item')
  assert set(out.row.keys()) == {k1, k2}
  src = out.node['item'][out.row[k1]]
  assert set(src.tolist()) == {0, 1} and out.col[k1].tolist() == [0, 0]
  got = set(zip(out.node['item'][out.row[k2]].tolist(), out.node['item'][out.col[k2]].tolist()))
  assert got == {(1, 0), (2, 1)}
  # per-type int64 tensors without trailing zero hops, like the reference (users only appear as seeds)
  assert out.num_sampled_nodes['user'].tolist() == [1] and out.num_sampled_nodes['item'].tolist() == [0, 2, 1]
  assert out.num_sampled_nodes['item'].size(0) == 3
  # in-direction: item seeds pull users
  ds_in = glt.data.Dataset(edge_dir='in')
  ds_in.init_graph({('user', 'u2i', 'item'): u2i, ('item', 'i2i', 'item'): i2i}, graph_mode='CPU')
  s = NeighborSampler(ds_in.graph, [2], edge_dir='in')
  out = s.sample_from_nodes(NodeSamplerInput(torch.tensor([1]), 'item'))
  k = ('user', 'u2i', 'item')
  assert set(out.node['user'][out.row[k]].tolist()) == {0, 1} and out.col[k].tolist() == [0, 0]


def test_hetero_id_tables_are_bounded_by_type_sizes():
  """Regression: the per-type id tables of a hetero sample used to be sized by the worst-case fan-out product
  (2^26 slots each, seconds per batch); they are now bounded by the number of nodes of the type."""
  import time
  from graphlearn_for_pytorch_b200.utils.synthetic import rmat_edges
  g = torch.Generator().manual_seed(0)
  nu, ni = 500, 300
  u2i = torch.stack([torch.randint(0, nu, (4000,), generator=g), torch.randint(0, ni, (4000,), generator=g)])
  ds = glt.data.Dataset(edge_dir='out')
  ds.init_graph({('u', 'to', 'i'): u2i, ('i', 'rev', 'u'): u2i.flip(0)}, graph_mode='CPU', num_nodes={'u': nu, 'i': ni})
  s = glt.sampler.NeighborSampler(ds.graph, [20, 20, 20, 20], device=torch.device('cpu'))
  assert s._hetero_table_cap(256) == 1 << 26                      # the unbounded estimate explodes ...
  assert s._hetero_type_bound('u', 256) <= nu + 256               # ... the per-type bound does not
  t0 = time.time()
  out = s.sample_from_nodes(glt.sampler.NodeSamplerInput(node=torch.arange(256), input_type='u'))
  assert time.time() - t0 < 5.0
  assert out.node['u'].numel() <= nu and out.node['i'].numel() <= ni


def test_behavioural_edge_cases_from_survey_appendix_c():
  """ids beyond the CSR get no neighbours, seeds are de-duplicated in first-occurrence order,
  `batch` is the seed prefix of `node`, fan-out -1 returns the whole neighbourhood, and a frontier that runs dry ends
  the expansion without failing."""
  ei = torch.tensor([[0, 0, 1, 2, 5], [1, 2, 2, 0, 0]])          # node 3, 4 isolated; node 5 -> 0; rows up to 5
  topo = glt.data.Topology(ei, layout='CSR')
  g = glt.data.Graph(topo, 'CPU', 0)
  s = glt.sampler.NeighborSampler(g, [-1, -1], device=torch.device('cpu'))
  out = s.sample_from_nodes(torch.tensor([2, 0, 2, 9]))           # duplicate seed, id 9 is outside the CSR
  assert out.node[:3].tolist() == [2, 0, 9] and out.batch.tolist() == [2, 0, 9]
  src, dst = out.node[out.col], out.node[out.row]                  # edge_dir='out': col = seed side, row = neighbour
  pairs = sorted(zip(src.tolist(), dst.tolist()))
  assert (2, 0) in pairs and (0, 1) in pairs and (0, 2) in pairs and all(a != 9 for a, _ in pairs)
  assert set(out.node.tolist()) == {0, 1, 2, 9}
  # a seed without out-edges: the expansion stops, the output is just the seed
  out = s.sample_from_nodes(torch.tensor([3]))
  assert out.node.tolist() == [3] and out.row.numel() == 0 and out.num_sampled_nodes[0] == 1
  assert sum(out.num_sampled_nodes) == 1 and sum(out.num_sampled_edges) == 0
  # one-hop API: neighbour counts are min(degree, fanout); ids beyond the CSR return 0
  nbr = glt.sampler.NeighborSampler(g, [1], device=torch.device('cpu')).sample_one_hop(torch.tensor([0, 3, 9]), 1)
  assert nbr.nbr_num.tolist() == [1, 0, 0]
