"""HeteroSageEngine (device-resident R-SAGE step) against a plain fp32 PyTorch recomputation on the SAME sampled
batch: forward loss, weight gradients of every (layer, type), and training progress under CUDA-graph replay."""
import os
import sys

import pytest
import torch

import graphlearn_for_pytorch_b200 as glt

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda', 0)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'examples'))


def _build(papers=4000, feat_dim=64, hidden=64, fan=(4, 3, 2), bs=256, use_graph=False, classes=19):
  from common import synthetic_igbh
  from graphlearn_for_pytorch_b200.models import HeteroSageEngine
  edges, feats, labels, sizes = synthetic_igbh(papers, papers // 2, 50, 30, feat_dim=feat_dim, num_classes=classes, seed=2)
  graphs, tables, keep = {}, {}, []
  for et, ei in edges.items():
    topo = glt.data.Topology(ei.to(DEV), layout='CSC', num_nodes=sizes[et[2]])
    graphs[et] = glt.data.Graph(topo, 'CUDA', 0)
  fb = {}
  for nt, x in feats.items():
    fb[nt] = x.to(DEV).to(torch.bfloat16)
    ut = glt.data.UnifiedTensor(0, torch.bfloat16)
    ut.append_shared_tensor(fb[nt])
    tables[nt] = ut._table()
    keep.append(ut)
  eng = HeteroSageEngine(graphs, tables, labels['paper'].to(DEV), feat_dim, sizes, 'paper', fanouts=list(fan),
                         batch_size=bs, hidden=hidden, num_classes=classes, edge_dir='in', lr=5e-3, seed=1,
                         device=DEV, use_cuda_graph=use_graph)
  eng._keep = keep
  return eng, fb, labels['paper'].to(DEV), sizes


def _reference(eng, fb, labels):
  """fp32 recomputation from the arena's COO, rounding to bf16 where the engine stores bf16."""
  nodes, rows, cols, _, nn, ne = eng.arena.to_coo()
  L, nt, kt = eng.L, eng.nt, eng.kt
  bf = lambda x: x.to(torch.bfloat16).float()
  W = {k: eng.p16[off:off + n * kk].view(n, kk).float().requires_grad_(True) for k, (off, n, kk) in eng._w_off.items()}
  B = {k: eng.p16[off:off + n].float() for k, (off, n) in eng._b_off.items()}
  x = {t: fb[eng.ntypes[t]][nodes[t]].float() for t in range(len(eng.ntypes))}
  for l in range(1, L + 1):
    nh = L - l + 1
    out = {}
    for t in eng.targets[l]:
      T = sum(nn[t][:nh])
      blocks = []
      for r in eng.in_rel[l][t]:
        E = sum(ne[r][:nh])
        src, dst = rows[r][:E], cols[r][:E]
        d = x[nt[r]].shape[1]
        agg = torch.zeros(T, d, device=DEV).index_add_(0, dst, x[nt[r]][src])
        deg = torch.zeros(T, device=DEV).index_add_(0, dst, torch.ones(E, device=DEV))
        blocks.append(bf(agg / deg.clamp(min=1).unsqueeze(1)))
      blocks.append(x[t][:T])
      A = torch.cat(blocks, 1)
      z = A @ W[(l, t)].t() + B[(l, t)]
      out[t] = bf(torch.relu(z)) if l < L else z
    x = out
  st = eng.tid[eng.seed_type]
  n0 = nn[st][0]
  logits = x[st][:n0, :eng.C]
  y = labels[nodes[st][:n0]]
  loss = torch.nn.functional.cross_entropy(logits, y)
  loss.backward()
  return float(loss), {k: w.grad for k, w in W.items()}, (nn, ne)


def test_hetero_engine_forward_backward_match_fp32():
  eng, fb, labels, sizes = _build()
  seeds = torch.randperm(sizes['paper'], device=DEV)[:eng.bs]
  eng._seeds.copy_(seeds)
  eng._sample()
  eng._forward()
  eng._backward()
  torch.cuda.synchronize()
  assert eng.overflow_count() == 0
  ref_loss, ref_grads, (nn, ne) = _reference(eng, fb, labels)
  assert abs(float(eng.loss.item()) - ref_loss) < 3e-2 * max(1.0, abs(ref_loss)), (float(eng.loss.item()), ref_loss)
  assert sum(sum(v) for v in ne) > 1000                     # the batch really has relations of every kind
  for k, (off, n, kk) in eng._w_off.items():
    g = eng.g32[off:off + n * kk].view(n, kk)
    r = ref_grads[k]
    if r is None:
      continue
    scale = max(float(r.abs().max()), 1e-6)
    err = float((g - r).abs().max()) / scale
    assert err < 6e-2, (k, err, scale)


def test_hetero_engine_trains_under_cuda_graph():
  eng, fb, labels, sizes = _build(use_graph=True)
  eng.warmup_and_capture(n_eager=1)
  assert eng._graph is not None
  g = torch.Generator().manual_seed(0)
  losses = []
  for i in range(60):
    seeds = torch.randperm(sizes['paper'], generator=g)[:eng.bs].to(DEV)
    losses.append(float(eng.train_step(seeds).item()))
  assert all(l == l for l in losses)
  assert sum(losses[-10:]) / 10 < sum(losses[:10]) / 10 - 0.05, (losses[:5], losses[-5:])
  l, c, n = eng.evaluate_batch(torch.arange(eng.bs, device=DEV))
  assert n == eng.bs and 0 <= c <= n
  nodes, edges = eng.batch_sizes()
  assert nodes['paper'][0] == eng.bs and sum(sum(v) for v in edges.values()) > 0
