"""Property-based tests (hypothesis) of the CPU operators on random graphs: invariants that must hold for ANY
input, complementing the hand-written fixtures (SURVEY 4.3 take-away (c)/(d))."""
import torch
from hypothesis import given, settings, strategies as st

import graphlearn_for_pytorch_b200 as glt
from graphlearn_for_pytorch_b200.ops import require_native

nat = require_native()


@st.composite
def graphs(draw, max_nodes=40, max_edges=200):
  n = draw(st.integers(2, max_nodes))
  e = draw(st.integers(0, max_edges))
  src = draw(st.lists(st.integers(0, n - 1), min_size=e, max_size=e))
  dst = draw(st.lists(st.integers(0, n - 1), min_size=e, max_size=e))
  return n, torch.tensor([src, dst], dtype=torch.int64).view(2, -1)


@settings(max_examples=60, deadline=None)
@given(graphs())
def test_csr_and_csc_are_permutations_of_the_edge_list(g):
  n, ei = g
  for layout in ('CSR', 'CSC'):
    topo = glt.data.Topology(ei, layout=layout, num_nodes=n)
    row, col, eids, _ = topo.to_coo()
    assert topo.indptr.numel() == n + 1 and int(topo.indptr[-1]) == ei.shape[1]
    assert torch.equal(ei[:, eids], torch.stack([row, col]))            # edge ids follow the permutation
    major = row if layout == 'CSR' else col
    assert torch.all(major[1:] >= major[:-1])                             # grouped by the major dimension


@settings(max_examples=60, deadline=None)
@given(graphs(), st.integers(1, 6), st.integers(0, 2 ** 31 - 1), st.booleans())
def test_one_hop_sampling_invariants(g, k, seed, replace):
  n, ei = g
  topo = glt.data.Topology(ei, layout='CSR', num_nodes=n)
  seeds = torch.arange(n)
  nbrs, cnt, eids = nat.cpu_sample_neighbors(topo.indptr, topo.indices, topo.edge_ids, seeds, k, True, replace, seed, 0)
  deg = topo.indptr[1:] - topo.indptr[:-1]
  assert torch.equal(cnt, torch.where(deg <= k, deg, torch.full_like(deg, k)) if not replace else
                     torch.where(deg == 0, deg, torch.where(deg <= k, deg, torch.full_like(deg, k))))
  off = 0
  for v in range(n):
    c = int(cnt[v])
    got, ge = nbrs[off:off + c], eids[off:off + c]
    truth = topo.indices[topo.indptr[v]:topo.indptr[v + 1]]
    assert all(int(x) in truth.tolist() for x in got)
    assert torch.equal(ei[1, ge], got) and torch.all(ei[0, ge] == v)     # returned edge ids are those edges
    if not replace:
      assert len(set(ge.tolist())) == c                                   # distinct edges without replacement
    off += c
  again = nat.cpu_sample_neighbors(topo.indptr, topo.indices, topo.edge_ids, seeds, k, True, replace, seed, 0)
  assert torch.equal(again[0], nbrs)                                      # pure function of (seed, stream, row)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.lists(st.integers(-1, 50), max_size=30), min_size=1, max_size=5))
def test_id_table_relabels_consistently(batches):
  t = nat.CpuIdTable(4)
  mapping = {}
  for b in batches:
    keys = torch.tensor(b, dtype=torch.int64)
    loc = t.insert(keys)
    for key, l in zip(b, loc.tolist()):
      if key < 0:
        assert l == -1
        continue
      assert mapping.setdefault(key, l) == l                              # stable across inserts
  inv = t.keys(0).tolist()
  assert len(inv) == len(mapping) == t.size()
  assert all(inv[l] == key for key, l in mapping.items())                 # keys() is the inverse map
  assert sorted(mapping.values()) == list(range(len(mapping)))            # dense, first-seen order


@settings(max_examples=40, deadline=None)
@given(graphs(max_nodes=30, max_edges=120), st.integers(1, 4), st.integers(0, 10 ** 6))
def test_multi_hop_output_is_a_consistent_subgraph(g, k, seed):
  n, ei = g
  ds = glt.data.Dataset()
  ds.init_graph(ei, graph_mode='CPU', num_nodes=n)
  s = glt.sampler.NeighborSampler(ds.graph, [k, k], device=torch.device('cpu'), seed=seed)
  seeds = torch.unique(torch.randint(0, n, (5,), generator=torch.Generator().manual_seed(seed)))
  out = s.sample_from_nodes(seeds)
  assert len(set(out.node.tolist())) == out.node.numel()                  # nodes are de-duplicated
  assert set(out.node[:seeds.numel()].tolist()) == set(seeds.tolist())    # seeds come first
  src, dst = out.node[out.col], out.node[out.row]                         # edge_dir='out': col = seed side
  edge_set = set(zip(ei[0].tolist(), ei[1].tolist()))
  assert all((a, b) in edge_set for a, b in zip(src.tolist(), dst.tolist()))
  assert sum(out.num_sampled_nodes) == out.node.numel() and sum(out.num_sampled_edges) == out.row.numel()


@settings(max_examples=15, deadline=None)
@given(graphs(max_nodes=30, max_edges=150), st.integers(2, 4), st.sampled_from(['by_src', 'by_dst']),
       st.integers(1, 9), st.booleans())
def test_partitioners_cover_every_node_and_edge_exactly_once(g, parts, strategy, chunk, use_range):
  """For ANY graph / partition count / edge-assignment strategy / chunk size: partitions are disjoint, their union
  is the graph, edges follow the chosen endpoint's owner, feature rows travel with their ids."""
  import tempfile
  from graphlearn_for_pytorch_b200.partition import RandomPartitioner, RangePartitioner, load_partition
  from graphlearn_for_pytorch_b200.utils.synthetic import id_features
  n, ei = g
  cls = RangePartitioner if use_range else RandomPartitioner
  with tempfile.TemporaryDirectory() as d:
    cls(d, parts, n, ei, node_feat=id_features(n, 4), edge_feat=id_features(max(ei.shape[1], 1), 2)[:ei.shape[1]],
        edge_assign_strategy=strategy, chunk_size=chunk).partition()
    nodes, edges = [], []
    for p in range(parts):
      _, _, gp, nf, ef, npb, epb = load_partition(d, p)
      rows, cols = gp.edge_index
      owner_side = rows if strategy == 'by_src' else cols
      assert torch.all(npb[owner_side] == p)
      assert torch.equal(ei[0][gp.eids], rows) and torch.equal(ei[1][gp.eids], cols)
      edges += gp.eids.tolist()
      assert torch.equal(nf.feats[:, 0].long(), nf.ids) and torch.all(npb[nf.ids] == p)
      nodes += nf.ids.tolist()
      if ef is not None and ef.ids.numel() > 0:
        assert torch.equal(ef.feats[:, 0].long(), ef.ids)
    assert sorted(nodes) == list(range(n)) and sorted(edges) == list(range(ei.shape[1]))
