"""Host-plane loaders and operators, this library vs the unmodified reference (baseline/_ref), same synthetic graph,
everything on the CPU (graph_mode='CPU', features in host memory).

  python benchmarks/bench_cpu_loaders.py ours | reference

Part 1 (500 k nodes / 10 M edges / F = 100): Dataset build, NeighborLoader [15,10,5] x 1024 seeds with features and
labels, LinkNeighborLoader with binary negatives, SubGraphLoader.  Part 2 (1 M / 20 M): Topology build (COO -> CSR /
CSC), sort_by_in_degree, Feature.cpu_get, negative sampling, sampling with edge ids, induced sub-graph, link sampling.
"""
import importlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
impl = sys.argv[1] if len(sys.argv) > 1 else 'ours'
scale = float(os.environ.get('SCALE', '1'))
if impl == 'reference':
  sys.path.insert(0, os.path.join(ROOT, 'baseline', 'shims'))
  sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))
  import graphlearn_torch as glt
else:
  sys.path.insert(0, ROOT)
  import graphlearn_for_pytorch_b200 as glt
cpu = torch.device('cpu')
res = {}


def graph(n, e, seed):
  g = torch.Generator().manual_seed(seed)
  src = torch.randint(0, n, (e,), generator=g)
  dst = (src + torch.randint(1, 5000, (e,), generator=g)) % n
  return g, torch.stack([src, dst])


def timed(name, fn, reps=1):
  fn()
  t = time.time()
  for _ in range(reps):
    out = fn()
  res[name] = round((time.time() - t) / reps, 4)
  return out


# ---------------------------------------------------------------- part 1: loaders
N, E = int(500_000 * scale), int(10_000_000 * scale)
g, ei = graph(N, E, 0)
x, y = torch.randn(N, 100, generator=g), torch.randint(0, 47, (N,), generator=g)
t = time.time()
ds = glt.data.Dataset()
ds.init_graph(ei, graph_mode='CPU', directed=True)
ds.init_node_features(x, with_gpu=False)
ds.init_node_labels(y)
res['dataset_build_s'] = round(time.time() - t, 3)
n_batches = max(int(100 * scale), 8)
seeds = torch.randperm(N, generator=g)[:n_batches * 1024]
loader = glt.loader.NeighborLoader(ds, [15, 10, 5], seeds, batch_size=1024, shuffle=True, drop_last=True, device=cpu)
for i, b in enumerate(loader):
  if i == 3:
    break
t, n, e = time.time(), 0, 0
for b in loader:
  n += 1
  e += b.edge_index.shape[1]
dt = time.time() - t
res['neighbor_loader_ms_per_batch'] = round(dt / n * 1e3, 2)
res['neighbor_loader_M_edges_per_s'] = round(e / dt / 1e6, 2)
ll = glt.loader.LinkNeighborLoader(ds, [10, 5], edge_label_index=ei[:, :max(n_batches // 2, 4) * 512],
                                   neg_sampling=glt.sampler.NegativeSampling('binary', 1), batch_size=512,
                                   shuffle=True, drop_last=True, device=cpu)
t, n = time.time(), 0
for b in ll:
  n += 1
res['link_loader_ms_per_batch'] = round((time.time() - t) / n * 1e3, 2)
SubGraphLoader = importlib.import_module(glt.__name__ + '.loader.subgraph_loader').SubGraphLoader
sl = SubGraphLoader(ds, seeds[:max(n_batches // 5, 4) * 256], [10, 5], batch_size=256, device=cpu)
t, n = time.time(), 0
for b in sl:
  n += 1
res['subgraph_loader_ms_per_batch'] = round((time.time() - t) / n * 1e3, 2)
del ds, loader, ll, sl

# ---------------------------------------------------------------- part 2: operators
N, E = int(1_000_000 * scale), int(20_000_000 * scale)
g, ei = graph(N, E, 1)
x = torch.randn(N, 64, generator=g)
topo = timed('topology_csr_s', lambda: glt.data.Topology(ei, input_layout='COO', layout='CSR'))
timed('topology_csc_s', lambda: glt.data.Topology(ei, input_layout='COO', layout='CSC'))
gr = glt.data.Graph(topo, 'CPU')
timed('sort_by_in_degree_s', lambda: glt.data.sort_by_in_degree(x, 0.0, topo))
f = glt.data.Feature(x, with_gpu=False)
ids = torch.randint(0, N, (200_000,), generator=g)
timed('feature_cpu_get_200k_s', lambda: f.cpu_get(ids), 5)
ns = glt.sampler.RandomNegativeSampler(gr, mode='CPU')
timed('negative_sample_100k_s', lambda: ns.sample(100_000, 5, True), 3)
s = glt.sampler.NeighborSampler(gr, [10, 5], device=cpu, with_edge=True)
sd = torch.randint(0, N, (1024,), generator=g)
timed('sample_with_edge_1024_s', lambda: s.sample_from_nodes(sd), 10)
timed('subgraph_1024_s', lambda: s.subgraph(glt.sampler.NodeSamplerInput(node=sd)), 3)
sl = glt.sampler.NeighborSampler(gr, [10, 5], device=cpu, with_neg=True)
es = glt.sampler.EdgeSamplerInput(row=ei[0, :1024].clone(), col=ei[1, :1024].clone(),
                                  neg_sampling=glt.sampler.NegativeSampling('binary', 1))
timed('sample_from_edges_1024_s', lambda: sl.sample_from_edges(es), 5)
print(json.dumps({'impl': impl, 'cpu_threads': torch.get_num_threads(), 'scale': scale, **res}))
