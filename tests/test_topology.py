import torch

import graphlearn_for_pytorch_b200 as glt


def test_coo_to_csr_and_back():
  row = torch.tensor([2, 0, 0, 1, 2, 2])
  col = torch.tensor([1, 2, 1, 0, 0, 3])
  topo = glt.data.Topology(torch.stack([row, col]), layout='CSR')
  assert topo.indptr.tolist() == [0, 2, 3, 6]
  assert topo.indices.tolist() == [1, 2, 0, 0, 1, 3]       # rows are column-sorted
  assert topo.edge_ids.tolist() == [2, 1, 3, 4, 0, 5]
  assert topo.degrees.tolist() == [2, 1, 3]
  r, c, e, _ = topo.to_coo()
  assert sorted(zip(r.tolist(), c.tolist())) == sorted(zip(row.tolist(), col.tolist()))
  assert topo.row_count == 3 and topo.edge_count == 6


def test_csc_layout_and_weights():
  row = torch.tensor([0, 0, 1, 2])
  col = torch.tensor([1, 2, 2, 0])
  w = torch.tensor([.1, .2, .3, .4])
  topo = glt.data.Topology(torch.stack([row, col]), edge_weights=w, layout='CSC')
  assert topo.indptr.tolist() == [0, 1, 2, 4]              # over columns
  assert topo.indices.tolist() == [2, 0, 0, 1]             # row ids
  assert torch.allclose(topo.edge_weights, torch.tensor([.4, .1, .2, .3]))
  ptr, ind, eid, ww = topo.to_csr()
  assert ptr.tolist() == [0, 2, 3, 4] and ind.tolist() == [1, 2, 2, 0]


def test_csr_input_passthrough():
  topo = glt.data.Topology((torch.tensor([0, 2, 3]), torch.tensor([1, 0, 0])), input_layout='CSR', layout='CSR')
  assert topo.indptr.tolist() == [0, 2, 3]
  topo2 = glt.data.Topology((torch.tensor([0, 2, 3]), torch.tensor([1, 0, 0])), input_layout='CSR', layout='CSC')
  r, c, _, _ = topo2.to_coo()
  assert sorted(zip(r.tolist(), c.tolist())) == [(0, 0), (0, 1), (1, 0)]


def test_typing_helpers():
  from graphlearn_for_pytorch_b200.typing import as_str, reverse_edge_type
  assert as_str(('a', 'r', 'b')) == 'a__r__b'
  assert reverse_edge_type(('a', 'r', 'b')) == ('b', 'rev_r', 'a')
  assert reverse_edge_type(('b', 'rev_r', 'a')) == ('a', 'r', 'b')
  assert reverse_edge_type(('a', 'r', 'a')) == ('a', 'r', 'a')
  assert glt.utils.parse_size('2GB') == 2 * 2 ** 30


def test_large_coo_to_csr_threaded_path_matches_stable_sort():
  """>= 2^20 edges take the row-range-parallel histogram / scatter of the native conversion: result must equal a
  stable (row, column) sort of the input, edge ids and weights included, with hub rows, empty rows and duplicates."""
  import torch
  from graphlearn_for_pytorch_b200.utils.topo import coo_to_csr
  g = torch.Generator().manual_seed(11)
  n, e = 50_000, (1 << 20) + 12_345
  rows = torch.randint(0, n, (e,), generator=g)
  rows[: e // 8] = 7                                       # a hub row
  rows[rows == 13] = 14                                    # an empty row
  cols = torch.randint(0, 1000, (e,), generator=g)         # many duplicate (row, col) pairs
  w = torch.rand(e, generator=g)
  eid = torch.randperm(e, generator=g)
  torch.set_num_threads(max(torch.get_num_threads(), 2))
  ptr, ind, oe, ow = coo_to_csr(rows, cols, eid, w, node_sizes=(n, 1000))
  perm = torch.argsort(cols, stable=True)
  perm = perm[torch.argsort(rows[perm], stable=True)]
  assert torch.equal(ptr[1:] - ptr[:-1], torch.bincount(rows, minlength=n)) and int(ptr[-1]) == e
  assert torch.equal(ind, cols[perm]) and torch.equal(oe, eid[perm]) and torch.equal(ow, w[perm])
  import pytest
  bad = rows.clone()
  bad[-1] = n
  with pytest.raises(RuntimeError):
    coo_to_csr(bad, cols, None, None, node_sizes=(n, 1000))
