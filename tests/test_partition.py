import tempfile

import torch

import graphlearn_for_pytorch_b200 as glt
from graphlearn_for_pytorch_b200.partition import (FrequencyPartitioner, RandomPartitioner, RangePartitionBook,
                                                    RangePartitioner, build_partition_feature,
                                                    cat_feature_cache, load_partition)
from graphlearn_for_pytorch_b200.utils.synthetic import id_features, ring_graph


def _check_homo(root, P, n, ei, with_feat=True, caching=False):
  seen_nodes, seen_edges = [], []
  for p in range(P):
    num, idx, g, nf, ef, npb, epb = load_partition(root, p, graph_caching=caching)
    assert num == P and idx == p
    rows, cols = g.edge_index
    if not caching:
      assert torch.all(npb[rows] == p)                       # edges follow their source owner
      assert torch.all(epb[g.eids] == p)
    assert torch.equal(ei[0][g.eids], rows) and torch.equal(ei[1][g.eids], cols)
    seen_edges += g.eids.tolist() if not caching else []
    if with_feat:
      assert torch.equal(nf.feats[:, 0].long(), nf.ids) and torch.all(npb[nf.ids] == p)
      assert torch.equal(ef.feats[:, 0].long(), ef.ids)
      seen_nodes += nf.ids.tolist()
  if with_feat:
    assert sorted(seen_nodes) == list(range(n))
  if not caching:
    assert sorted(seen_edges) == list(range(ei.shape[1]))


def test_random_and_range_partitioner_homo():
  n = 40
  ei = ring_graph(n)
  for cls in (RandomPartitioner, RangePartitioner):
    for caching in (False, True):
      with tempfile.TemporaryDirectory() as d:
        cls(d, 2, n, ei, node_feat=id_features(n, 8), edge_feat=id_features(ei.shape[1], 4), chunk_size=7) \
          .partition(graph_caching=caching)
        _check_homo(d, 2, n, ei, caching=caching)


def test_two_stage_feature_build():
  n = 40
  ei = ring_graph(n)
  with tempfile.TemporaryDirectory() as d:
    RandomPartitioner(d, 2, n, ei, chunk_size=9).partition(with_feature=False)
    for p in range(2):
      build_partition_feature(d, p, chunk_size=5, node_feat=id_features(n, 8), edge_feat=id_features(ei.shape[1], 4))
    _check_homo(d, 2, n, ei)


def test_random_partitioner_hetero():
  u2i = torch.tensor([[0, 0, 1, 2, 3, 3], [0, 1, 1, 2, 3, 0]])
  i2i = torch.tensor([[0, 1, 2, 3], [1, 2, 3, 0]])
  with tempfile.TemporaryDirectory() as d:
    RandomPartitioner(d, 2, {'user': 4, 'item': 4}, {('user', 'u2i', 'item'): u2i, ('item', 'i2i', 'item'): i2i},
                      node_feat={'user': id_features(4, 4), 'item': id_features(4, 4)}).partition()
    got = {('user', 'u2i', 'item'): 0, ('item', 'i2i', 'item'): 0}
    for p in range(2):
      num, idx, g, nf, ef, npb, epb = load_partition(d, p)
      for et, gp in g.items():
        assert torch.all(npb[et[0]][gp.edge_index[0]] == p)
        got[et] += gp.eids.numel()
      for nt, f in nf.items():
        assert torch.equal(f.feats[:, 0].long(), f.ids)
    assert got[('user', 'u2i', 'item')] == 6 and got[('item', 'i2i', 'item')] == 4


def test_frequency_partitioner_and_cache():
  n = 40
  ei = ring_graph(n)
  probs = [torch.zeros(n), torch.zeros(n)]
  probs[0][:20] = 1.0; probs[0][20:24] = 0.5
  probs[1][20:] = 1.0; probs[1][:3] = 0.7
  with tempfile.TemporaryDirectory() as d:
    FrequencyPartitioner(d, 2, n, ei, probs, node_feat=id_features(n, 8), cache_ratio=0.1, chunk_size=40).partition()
    num, _, g, nf, _, npb, _ = load_partition(d, 0)
    assert int((npb[:20] == 0).sum()) >= 16 and int((npb[20:] == 1).sum()) >= 16
    assert nf.cache_ids is not None and torch.all(npb[nf.cache_ids] == 1)       # hot *remote* rows
    assert torch.equal(nf.cache_feats[:, 0].long(), nf.cache_ids)
    ratio, feats, id2index, new_pb = cat_feature_cache(0, nf, npb)
    assert ratio > 0 and torch.all(new_pb[nf.cache_ids] == 0)
    ids = torch.cat([nf.ids, nf.cache_ids])
    assert torch.equal(feats[id2index[ids]][:, 0].long(), ids)


def test_range_partition_book():
  pb = RangePartitionBook([(0, 10), (10, 25), (25, 40)], 1)
  assert pb[torch.tensor([0, 9, 10, 24, 25, 39])].tolist() == [0, 0, 1, 1, 2, 2]
  assert pb.id2index[torch.tensor([10, 12])].tolist() == [0, 2]
  assert pb.id_filter(pb, 2).tolist() == list(range(25, 40))
  assert pb.bounds == [0, 10, 25, 40]


def test_frequency_chunk_assignment_respects_quota_and_follows_hotness():
  """The vectorised chunk assignment gives every node an owner, never exceeds the per-chunk quota, and agrees with
  the sequential "best (partition, node) pairs first" walk it replaced (ties aside)."""
  from graphlearn_for_pytorch_b200.partition.frequency_partitioner import _assign_chunk

  def sequential(score, quota):
    P, c = score.shape
    owner, load = [-1] * c, [0] * P
    for flat in torch.argsort(score.flatten(), descending=True).tolist():
      p, v = divmod(flat, c)
      if owner[v] < 0 and load[p] < quota:
        owner[v], load[p] = p, load[p] + 1
    return torch.tensor(owner)
  g = torch.Generator().manual_seed(3)
  for P, c in ((4, 2000), (7, 701), (2, 5)):
    probs = torch.rand(P, c, generator=g) ** 3
    score = P * probs - probs.sum(0, keepdim=True)
    quota = (c + P - 1) // P
    own = _assign_chunk(score, quota)
    assert int(own.min()) >= 0 and int(torch.bincount(own, minlength=P).max()) <= quota
    ref = sequential(score, quota)
    assert float((own == ref).float().mean()) > 0.97
    mass, ref_mass = probs[own, torch.arange(c)].sum(), probs[ref, torch.arange(c)].sum()
    assert float(mass) >= 0.995 * float(ref_mass)
  cold = _assign_chunk(torch.zeros(4, 10), 3)                 # no hotness information: still balanced
  assert int(cold.min()) >= 0 and int(torch.bincount(cold, minlength=4).max()) <= 3


def test_partition_benchmark_runs_small():
  import os
  import subprocess
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  for kind in ('random', 'frequency'):
    out = subprocess.run([sys.executable, 'benchmarks/bench_partition_cpu.py', 'ours', kind], cwd=root,
                         capture_output=True, text=True, timeout=300, env=dict(os.environ, N='20000', E='200000'))
    assert out.returncode == 0 and '"M_edges_per_s"' in out.stdout, out.stderr[-2000:]
