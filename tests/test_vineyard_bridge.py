"""Property-graph fragment bridge (reference: data/vineyard_utils.py + v6d/vineyard_utils.cc; the
reference has no test for it because it needs a vineyard daemon -- here the Arrow fragment store
stands in for the daemon, see data/vineyard_utils.py)."""
import torch

import graphlearn_for_pytorch_b200 as glt
from graphlearn_for_pytorch_b200.data import vineyard_utils as vu


def _graph(n=200, e=1500, seed=0):
  g = torch.Generator().manual_seed(seed)
  src = torch.randint(0, n, (e,), generator=g)
  dst = torch.randint(0, n, (e,), generator=g)
  feat = torch.arange(n, dtype=torch.float32).unsqueeze(1).repeat(1, 8)
  label = torch.arange(n) % 7
  w = torch.rand(e, generator=g)
  return src, dst, feat, label, w


def test_fragment_csr_features_and_books(tmp_path):
  n = 200
  src, dst, feat, label, w = _graph(n)
  ids = vu.write_arrow_fragments(str(tmp_path), 'g', 2, {'v': {'feat': feat, 'label': label, 'f0': feat[:, 0]}},
                                 {'e': ('v', 'v', torch.stack([src, dst]), {'w': w, 'ef': torch.stack([src, dst], 1).float()})})
  assert ids == ['g_0', 'g_1']
  seen = 0
  for f, oid in enumerate(ids):
    off = vu.get_frag_vertex_offset(str(tmp_path), oid, 'v')
    num = vu.get_frag_vertex_num(str(tmp_path), oid, 'v')
    assert (off, num) == (f * 100, 100)
    indptr, indices, eids = vu.vineyard_to_csr(str(tmp_path), oid, 'v', 'e', 'out', True)
    assert indptr.numel() == num + 1 and indices.numel() == eids.numel() == int(((src >= off) & (src < off + num)).sum())
    seen += indices.numel()
    ef = vu.load_edge_feature_from_vineyard(str(tmp_path), oid, ['ef'], 'e')
    for v in (0, 17, 99):
      nbrs = indices[indptr[v]:indptr[v + 1]]
      assert sorted(nbrs.tolist()) == sorted(dst[src == off + v].tolist())
      rows = ef[eids[indptr[v]:indptr[v + 1]]]     # edge features follow the returned edge ids
      assert torch.equal(rows[:, 0].long(), torch.full_like(nbrs, off + v)) and torch.equal(rows[:, 1].long(), nbrs)
    x = vu.load_vertex_feature_from_vineyard(str(tmp_path), oid, ['feat', 'f0'], 'v')
    assert x.shape == (num, 9) and torch.equal(x[:, 0], torch.arange(off, off + num).float())
    g2l = vu.VineyardGid2Lid(str(tmp_path), oid, 'v')
    assert len(g2l) == num and int(g2l[torch.tensor([off + 5])][0]) == 5
  assert seen == src.numel()
  pb = vu.VineyardPartitionBook(str(tmp_path), 'g_1', 'v')
  assert pb[torch.tensor([0, 99, 100, 199])].tolist() == [0, 0, 1, 1] and pb.offset == 100
  assert torch.equal(vu.v6d_id_filter(pb, 1), torch.arange(100, 200))
  assert vu.v6d_id_select(torch.tensor([100, 3, 150]), torch.tensor([True, False, True]), pb).tolist() == [0, 50]
  assert vu.get_fid_from_gid(torch.tensor([3, 150]), str(tmp_path), 'g_0', 'v').tolist() == [0, 1]
  remap = vu.VineyardPartitionBook(str(tmp_path), 'g_0', 'v', fid2pid={0: 1, 1: 0})
  assert remap[torch.tensor([0, 150])].tolist() == [1, 0]


def test_dataset_load_vineyard_homo_and_sampling(tmp_path):
  n = 200
  src, dst, feat, label, w = _graph(n)
  vu.write_arrow_fragments(str(tmp_path), 'g', 1, {'v': {'feat': feat, 'label': label}},
                           {'e': ('v', 'v', torch.stack([src, dst]), {'w': w})})
  ds = glt.data.Dataset(edge_dir='out')
  ds.load_vineyard('g_0', str(tmp_path), [('v', 'e', 'v')], edge_weights={('v', 'e', 'v'): 'w'},
                   node_features={'v': ['feat']}, node_labels={'v': 'label'})
  topo = ds.graph.topo
  assert topo.edge_count == src.numel() and torch.equal(ds.node_labels, label)
  # weights follow the CSR order
  for v in (1, 50):
    sl = slice(int(topo.indptr[v]), int(topo.indptr[v + 1]))
    got = sorted(zip(topo.indices[sl].tolist(), [round(x, 5) for x in topo.edge_weights[sl].tolist()]))
    exp = sorted(zip(dst[src == v].tolist(), [round(x, 5) for x in w[src == v].tolist()]))
    assert got == exp
  loader = glt.loader.NeighborLoader(ds, [3, 2], torch.arange(40), batch_size=20, shuffle=False, device=torch.device('cpu'))
  nb = 0
  for b in loader:
    assert torch.equal(b.x[:, 0].long(), b.node) and torch.equal(b.y, b.node % 7)
    nb += 1
  assert nb == 2


def test_dataset_load_vineyard_hetero_in_edges(tmp_path):
  nu, ni = 60, 40
  g = torch.Generator().manual_seed(1)
  u = torch.randint(0, nu, (500,), generator=g)
  i = torch.randint(0, ni, (500,), generator=g)
  vu.write_arrow_fragments(str(tmp_path), 'h', 1,
                           {'user': {'x': torch.arange(nu).float().unsqueeze(1).repeat(1, 4)},
                            'item': {'x': torch.arange(ni).float().unsqueeze(1).repeat(1, 4), 'y': torch.arange(ni) % 3}},
                           {'buys': ('user', 'item', torch.stack([u, i]), {})}, edge_owner='dst')
  ds = glt.data.Dataset(edge_dir='in')
  et = ('user', 'buys', 'item')
  ds.load_vineyard('h_0', str(tmp_path), [et], node_features={'user': ['x'], 'item': ['x']}, node_labels={'item': 'y'})
  topo = ds.graph[et].topo                      # CSC keyed by item
  for v in (0, 7, 39):
    assert sorted(topo.indices[topo.indptr[v]:topo.indptr[v + 1]].tolist()) == sorted(u[i == v].tolist())
  assert torch.equal(ds.node_labels['item'], torch.arange(ni) % 3)
  assert torch.equal(ds.node_features['user'].cpu_get(torch.tensor([5, 59]))[:, 0], torch.tensor([5., 59.]))


def test_dist_dataset_load_vineyard_global_keys(tmp_path):
  n = 200
  src, dst, feat, label, w = _graph(n)
  vu.write_arrow_fragments(str(tmp_path), 'g', 2, {'v': {'feat': feat, 'label': label}},
                           {'e': ('v', 'v', torch.stack([src, dst]), {})})
  import graphlearn_for_pytorch_b200.distributed as gd
  ds = gd.DistDataset(edge_dir='out')
  ds.load_vineyard('g_1', str(tmp_path), [('v', 'e', 'v')], node_features={'v': ['feat']}, node_labels={'v': 'label'})
  assert (ds.num_partitions, ds.partition_idx) == (2, 1)
  topo = ds.graph.topo
  assert topo.indptr.numel() == n + 1 and int(topo.indptr[100]) == 0       # rows of fragment 0 are empty
  assert sorted(topo.indices[topo.indptr[150]:topo.indptr[151]].tolist()) == sorted(dst[src == 150].tolist())
  assert ds.node_pb[torch.tensor([5, 150])].tolist() == [0, 1]
  assert ds.node_labels[150] == 150 % 7 and ds.node_labels[5] == -1
  assert torch.equal(ds.node_features.cpu_get(torch.tensor([150, 199]))[:, 0], torch.tensor([150., 199.]))
  ds.random_node_split(0.1, 0.1)
  allidx = torch.cat([ds.train_idx, ds.val_idx, ds.test_idx])
  assert allidx.min() >= 100 and allidx.numel() == 100
