import torch

import graphlearn_for_pytorch_b200 as glt
from helpers import adjacency_sets, rmat_csr


def test_one_hop_membership_and_counts(native):
  _, topo = rmat_csr()
  adj = adjacency_sets(topo)
  seeds = torch.arange(0, 2000, 7)
  for k in (1, 3, 10):
    nbr, cnt, eid = native.cpu_sample_neighbors(topo.indptr, topo.indices, topo.edge_ids, seeds, k, True,
                                                False, 123, 0)
    off = 0
    for s, c in zip(seeds.tolist(), cnt.tolist()):
      deg = int(topo.degrees[s])
      assert c == min(deg, k)
      got = nbr[off:off + c].tolist()
      assert set(got) <= adj[s]
      # without replacement: picked *positions* are distinct
      assert len(set(eid[off:off + c].tolist())) == c
      off += c
    assert off == nbr.numel()


def test_all_neighbors_and_out_of_range(native):
  _, topo = rmat_csr(200, 2000)
  seeds = torch.tensor([0, 5, 199, 100000])
  nbr, cnt, _ = native.cpu_sample_neighbors(topo.indptr, topo.indices, None, seeds, -1, False, False, 1, 0)
  assert cnt[-1] == 0
  assert cnt[:3].tolist() == topo.degrees[seeds[:3]].tolist()


def test_determinism_and_stream_independence(native):
  _, topo = rmat_csr()
  seeds = torch.arange(100)
  a = native.cpu_sample_neighbors(topo.indptr, topo.indices, None, seeds, 4, False, False, 9, 3)[0]
  b = native.cpu_sample_neighbors(topo.indptr, topo.indices, None, seeds, 4, False, False, 9, 3)[0]
  c = native.cpu_sample_neighbors(topo.indptr, topo.indices, None, seeds, 4, False, False, 9, 4)[0]
  assert torch.equal(a, b) and not torch.equal(a, c)


def test_uniformity(native):
  # one row with 20 neighbours, k=5: each neighbour should be picked w.p. 1/4
  indptr = torch.tensor([0, 20])
  indices = torch.arange(100, 120)
  hits = torch.zeros(20)
  trials = 4000
  for t in range(trials):
    nbr, _, _ = native.cpu_sample_neighbors(indptr, indices, None, torch.tensor([0]), 5, False, False, 77, t)
    hits[nbr - 100] += 1
  p = hits / trials
  assert (p - 0.25).abs().max() < 0.04


def test_weighted_sampling_statistics(native):
  indptr = torch.tensor([0, 4])
  indices = torch.tensor([10, 11, 12, 13])
  w = torch.tensor([1.0, 1.0, 2.0, 4.0])
  hits = torch.zeros(4)
  trials = 4000
  for t in range(trials):
    nbr, cnt, _ = native.cpu_sample_neighbors_weighted(indptr, indices, None, w, torch.tensor([0]), 1, False, 5, t)
    assert cnt.item() == 1
    hits[nbr - 10] += 1
  p = hits / trials
  assert (p - w / w.sum()).abs().max() < 0.04


def test_negative_sampler(native):
  ei, topo = rmat_csr(300, 6000)
  edges = set(zip(ei[0].tolist(), ei[1].tolist()))
  rows, cols = native.cpu_negative_sample(topo.indptr, topo.indices, 300, 300, 500, 5, False, True, 3, 0)
  assert rows.numel() <= 500 and rows.numel() > 400
  assert all((r, c) not in edges for r, c in zip(rows.tolist(), cols.tolist()))
  rows, cols = native.cpu_negative_sample(topo.indptr, topo.indices, 300, 300, 500, 1, True, True, 3, 1)
  assert rows.numel() == 500


def test_subgraph_op(native):
  # 0->1, 0->2, 1->2, 2->0, 3->0
  topo = glt.data.Topology(torch.tensor([[0, 0, 1, 2, 3], [1, 2, 2, 0, 0]]), layout='CSR', num_nodes=4)
  nodes, rows, cols, eids = native.cpu_node_subgraph(topo.indptr, topo.indices, topo.edge_ids,
                                                     torch.tensor([2, 0, 2]), True)
  assert nodes.tolist() == [2, 0]
  got = set(zip(nodes[rows].tolist(), nodes[cols].tolist()))
  assert got == {(0, 2), (2, 0)}
  assert sorted(eids.tolist()) == [1, 3]


def test_id_table(native):
  t = native.CpuIdTable(4)
  ids = t.insert(torch.tensor([7, 3, 7, 9, -1]))
  assert ids.tolist() == [0, 1, 0, 2, -1]
  assert t.size() == 3 and t.keys().tolist() == [7, 3, 9]
  for i in range(100):  # growth
    t.insert(torch.tensor([1000 + i]))
  assert t.size() == 103
  assert t.lookup(torch.tensor([9, 1050, 5])).tolist() == [2, 53, -1]
  assert t.keys(100).tolist() == [1097, 1098, 1099]


def test_random_walk(native):
  _, topo = rmat_csr(500, 8000)
  adj = adjacency_sets(topo)
  starts = torch.arange(0, 500, 5)
  for p, q in ((1.0, 1.0), (0.5, 2.0)):
    walks = native.cpu_random_walk(topo.indptr, topo.indices, starts, 6, p, q, 11, 0)
    assert walks.shape == (100, 7)
    for w in walks.tolist():
      for a, b in zip(w[:-1], w[1:]):
        assert b in adj[a] or (len(adj[a]) == 0 and a == b)


def test_stitch(native):
  idx = [torch.tensor([0, 2]), torch.tensor([1])]
  nbrs = [torch.tensor([10, 11, 12]), torch.tensor([20])]
  nums = [torch.tensor([1, 2]), torch.tensor([1])]
  eids = [torch.tensor([1, 2, 3]), torch.tensor([9])]
  n, c, e = native.cpu_stitch(3, idx, nbrs, nums, eids)
  assert n.tolist() == [10, 20, 11, 12] and c.tolist() == [1, 1, 2] and e.tolist() == [1, 9, 2, 3]
