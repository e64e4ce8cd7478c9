"""`py_graphlearn_torch` facade: the native-module names of the reference (python/py_export_glt.cc:47-222) on top of
this package's operators.  CPU classes are checked here; the CUDA classes share the code path and are covered by a
gpu-marked case."""
import pickle

import pytest
import torch

import graphlearn_for_pytorch_b200 as glt

pywrap = glt.py_graphlearn_torch

INDPTR = torch.tensor([0, 2, 4, 6, 8, 9, 10])
INDICES = torch.tensor([1, 2, 2, 3, 3, 4, 4, 5, 5, 0])
EIDS = torch.arange(10)


def _graph(weights=None):
  g = pywrap.Graph()
  g.init_cpu_from_csr(INDPTR, INDICES, EIDS, weights)
  return g


def _adj():
  return {r: INDICES[INDPTR[r]:INDPTR[r + 1]].tolist() for r in range(6)}


def test_handles_resolve_to_one_module():
  assert glt.data.pywrap is pywrap and glt.sampler.pywrap is pywrap
  names = ['Graph', 'GraphMode', 'SubGraph', 'CPURandomSampler', 'CPUWeightedSampler', 'CUDARandomSampler',
           'CPURandomNegativeSampler', 'CUDARandomNegativeSampler', 'CPUInducer', 'CUDAInducer', 'CPUHeteroInducer',
           'CUDAHeteroInducer', 'CPUSubGraphOp', 'CUDASubGraphOp', 'SampleQueue', 'QueueTimeoutError', 'UnifiedTensor',
           'SharedTensor', 'RandomSeedManager', 'cpu_stitch_sample_results', 'cuda_stitch_sample_results']
  assert not [n for n in names if not hasattr(pywrap, n)]


def test_graph_and_random_sampler():
  g = _graph()
  assert (g.get_row_count(), g.get_col_count(), g.get_edge_count()) == (6, 6, 10)
  ids = torch.tensor([0, 1, 4, 3])
  nbr, num, eid = pywrap.CPURandomSampler(g).sample_with_edge(ids, 2)
  assert num.tolist() == [2, 2, 1, 2]
  adj, off = _adj(), 0
  for r, n in zip(ids.tolist(), num.tolist()):
    for j in range(off, off + n):
      assert nbr[j].item() in adj[r]
      assert INDICES[eid[j]].item() == nbr[j].item()      # the edge id names the sampled edge
    off += n
  nbr2, num2 = pywrap.CPURandomSampler(g).sample(ids, 1)
  assert num2.tolist() == [1, 1, 1, 1] and nbr2.numel() == 4


def test_weighted_sampler_follows_weights():
  w = torch.tensor([1., 0., 0., 1., 1., 0., 0., 1., 1., 1.])     # one live edge per row
  nbr, num = pywrap.CPUWeightedSampler(_graph(w)).sample(torch.tensor([0, 1, 2, 3]), 1)
  assert num.tolist() == [1, 1, 1, 1] and nbr.tolist() == [1, 3, 3, 5]


def test_inducer_relabels_incrementally():
  g = _graph()
  ind = pywrap.CPUInducer(16)
  seeds = ind.init_node(torch.tensor([0, 1, 4, 0]))
  assert seeds.tolist() == [0, 1, 4]
  nbr, num = pywrap.CPURandomSampler(g).sample(seeds, 2)
  nodes, rows, cols = ind.induce_next(seeds, nbr, num)
  known = seeds.tolist() + nodes.tolist()
  assert len(set(known)) == len(known)                           # only nodes unseen so far are reported
  assert rows.tolist() == [0, 0, 1, 1, 2]
  assert [known[c] for c in cols.tolist()] == nbr.tolist()
  # a second hop continues the numbering
  nbr2, num2 = pywrap.CPURandomSampler(g).sample(nodes, 2)
  nodes2, rows2, cols2 = ind.induce_next(nodes, nbr2, num2)
  known2 = known + nodes2.tolist()
  assert [known2[c] for c in cols2.tolist()] == nbr2.tolist()
  assert rows2.min().item() >= len(seeds)


def test_hetero_inducer():
  hi = pywrap.CPUHeteroInducer({'a': 16, 'b': 16})
  assert hi.init_node({'a': torch.tensor([0, 1, 0])})['a'].tolist() == [0, 1]
  nodes, rows, cols = hi.induce_next({
    ('a', 'r', 'b'): (torch.tensor([0, 1]), torch.tensor([3, 4, 4]), torch.tensor([2, 1])),
    ('a', 's', 'a'): (torch.tensor([0, 1]), torch.tensor([1, 5]), torch.tensor([1, 1]))})
  assert nodes['b'].tolist() == [3, 4] and nodes['a'].tolist() == [5]
  assert rows[('a', 'r', 'b')].tolist() == [0, 0, 1] and cols[('a', 'r', 'b')].tolist() == [0, 1, 1]
  assert rows[('a', 's', 'a')].tolist() == [0, 1] and cols[('a', 's', 'a')].tolist() == [1, 2]


def test_subgraph_op_reports_stored_orientation():
  sg = pywrap.CPUSubGraphOp(_graph()).node_subgraph(torch.tensor([3, 0, 1, 2]), True)
  assert sg.nodes.tolist() == [0, 1, 2, 3]
  got = sorted(zip(sg.nodes[sg.rows].tolist(), sg.nodes[sg.cols].tolist(), sg.eids.tolist()))
  want = sorted((r, c, int(EIDS[INDPTR[r] + k])) for r in range(4) for k, c in enumerate(_adj()[r]) if c < 4)
  assert got == want


def test_negative_sampler_avoids_edges():
  rows, cols = pywrap.CPURandomNegativeSampler(_graph()).sample(8, 5, False)
  adj = _adj()
  assert rows.numel() == cols.numel() <= 8
  assert all(c not in adj[r] for r, c in zip(rows.tolist(), cols.tolist()))


def test_sample_queue_roundtrip_and_pickle():
  q = pywrap.SampleQueue(4, 1 << 20)
  assert q.empty()
  q.send({'ids': torch.arange(5), 'x': torch.ones(2, 3)})
  q2 = pickle.loads(pickle.dumps(q))                             # attaches to the same shared-memory ring
  msg = q2.receive(1000)
  assert msg['ids'].tolist() == [0, 1, 2, 3, 4] and msg['x'].shape == (2, 3)
  with pytest.raises(pywrap.QueueTimeoutError):
    q.receive(20)


def test_seed_manager_and_stitch():
  pywrap.RandomSeedManager.getInstance().setSeed(7)
  assert pywrap.RandomSeedManager.getInstance().getSeed() == 7
  g = _graph()
  ids = torch.tensor([0, 1, 2, 3, 4])
  a = pywrap.CPURandomSampler(g).sample(ids, 2)
  pywrap.RandomSeedManager.getInstance().setSeed(7)
  b = pywrap.CPURandomSampler(g).sample(ids, 2)
  assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
  nbrs, num, eids = pywrap.cpu_stitch_sample_results(
    torch.tensor([5, 6, 7]), [torch.tensor([0, 2]), torch.tensor([1])], [torch.tensor([1, 2, 3]), torch.tensor([9])],
    [torch.tensor([2, 1]), torch.tensor([1])], [])
  assert nbrs.tolist() == [1, 2, 9, 3] and num.tolist() == [2, 1, 1] and eids is None


@pytest.mark.gpu
def test_cuda_classes():
  dev = torch.device('cuda:0')
  g = pywrap.Graph()
  g.init_cuda_from_csr(INDPTR, INDICES, 0, pywrap.GraphMode.DMA, EIDS)
  assert g.get_mode() == pywrap.GraphMode.DMA and g.get_edge_count() == 10
  ids = torch.tensor([0, 1, 4], device=dev)
  nbr, num, eid = pywrap.CUDARandomSampler(g).sample_with_edge(ids, 2)
  assert nbr.is_cuda and num.tolist() == [2, 2, 1]
  assert torch.equal(INDICES.to(dev)[eid], nbr)
  ind = pywrap.CUDAInducer(16)
  seeds = ind.init_node(ids)
  nodes, rows, cols = ind.induce_next(seeds, nbr, num)
  known = torch.cat([seeds, nodes])
  assert torch.equal(known[cols], nbr) and rows.tolist() == [0, 0, 1, 1, 2]
  sg = pywrap.CUDASubGraphOp(g).node_subgraph(torch.tensor([0, 1, 2, 3], device=dev), True)
  assert sorted(zip(sg.nodes[sg.rows].tolist(), sg.nodes[sg.cols].tolist())) == \
      [(0, 1), (0, 2), (1, 2), (1, 3), (2, 3)]
  rows, cols = pywrap.CUDARandomNegativeSampler(g).sample(8, 5, False)
  assert rows.is_cuda and rows.numel() == cols.numel()
  nb, nn, _ = pywrap.cuda_stitch_sample_results(
    torch.tensor([5, 6, 7], device=dev), [torch.tensor([0, 2], device=dev), torch.tensor([1], device=dev)],
    [torch.tensor([1, 2, 3], device=dev), torch.tensor([9], device=dev)],
    [torch.tensor([2, 1], device=dev), torch.tensor([1], device=dev)], [])
  assert nb.tolist() == [1, 2, 9, 3] and nn.tolist() == [2, 1, 1]


def test_inducer_matches_the_reference_native_contract():
  """Inputs and expected outputs of the reference's own native test (test/cpp/test_inducer.cu:56-100): first-seen
  numbering, only unseen nodes reported per hop, one source entry per neighbour-count entry."""
  srcs1, nbrs1, num1 = torch.tensor([0, 1, 2, 2, 3]), torch.tensor([1, 2, 2, 3, 4, 5, 4, 5, 1]), torch.tensor([2, 2, 2, 2, 1])
  srcs2, nbrs2, num2 = torch.tensor([0, 1, 2, 3, 4, 5]), torch.tensor([1, 7, 2, 3, 4, 6, 7]), torch.tensor([1, 1, 1, 2, 1, 1])
  ind = pywrap.CPUInducer(srcs1.numel() + nbrs1.numel() + srcs2.numel() + nbrs2.numel())
  assert ind.init_node(srcs1).tolist() == [0, 1, 2, 3]
  nodes, rows, cols = ind.induce_next(srcs1, nbrs1, num1)
  assert (nodes.tolist(), rows.tolist(), cols.tolist()) == \
      ([4, 5], [0, 0, 1, 1, 2, 2, 2, 2, 3], [1, 2, 2, 3, 4, 5, 4, 5, 1])
  nodes, rows, cols = ind.induce_next(srcs2, nbrs2, num2)
  assert (nodes.tolist(), rows.tolist(), cols.tolist()) == ([7, 6], [0, 1, 2, 3, 3, 4, 5], [1, 6, 2, 3, 4, 7, 6])


def test_hetero_inducer_matches_the_reference_native_contract():
  """test/cpp/test_hetero_inducer.cu:30-80 (meta-path a -> b, a -> c)."""
  hi = pywrap.CPUHeteroInducer({'a': 10, 'b': 10, 'c': 10})
  assert hi.init_node({'a': torch.tensor([0, 1, 2, 2, 3])})['a'].tolist() == [0, 1, 2, 3]
  a2b, a2c = ('a', 'a2b', 'b'), ('a', 'a2c', 'c')
  nodes, rows, cols = hi.induce_next({
    a2b: (torch.tensor([0, 1, 2, 2, 3]), torch.tensor([1, 2, 2, 3, 4, 5, 4, 5, 1]), torch.tensor([2, 2, 2, 2, 1])),
    a2c: (torch.tensor([2, 1, 3, 2, 3]), torch.tensor([3, 5, 2, 3, 4, 3, 1, 2, 1]), torch.tensor([1, 2, 2, 3, 1]))})
  assert nodes['b'].tolist() == [1, 2, 3, 4, 5] and nodes['c'].tolist() == [3, 5, 2, 4, 1] and 'a' not in nodes
  assert rows[a2b].tolist() == [0, 0, 1, 1, 2, 2, 2, 2, 3] and cols[a2b].tolist() == [0, 1, 1, 2, 3, 4, 3, 4, 0]
  assert rows[a2c].tolist() == [2, 1, 1, 3, 3, 2, 2, 2, 3] and cols[a2c].tolist() == [0, 1, 2, 0, 3, 0, 4, 2, 4]


def test_stitch_matches_the_reference_native_contract():
  """test/cpp/test_stitch_sample_results.cu:26-60: three partitions answer for interleaved positions of six seeds."""
  T = torch.tensor
  nbrs, num, eids = pywrap.cpu_stitch_sample_results(
    T([1, 2, 3, 4, 5, 6]), [T([0, 2]), T([1, 5]), T([3, 4])],
    [T([4, 5]), T([3, 7, 8, 9, 10, 11]), T([5, 6, 7, 6, 7, 8, 9])],
    [T([0, 2]), T([1, 5]), T([3, 4])],
    [T([1, 2]), T([0, 10, 11, 12, 13, 14]), T([3, 4, 5, 6, 7, 8, 9])])
  assert num.tolist() == [0, 1, 2, 3, 4, 5]
  assert nbrs.tolist() == [3, 4, 5, 5, 6, 7, 6, 7, 8, 9, 7, 8, 9, 10, 11]
  assert eids.tolist() == [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14]
