"""CPU-mode NeighborSampler throughput (graph_mode='CPU', device=cpu): this library vs the unmodified reference
(baseline/_ref) on the same synthetic graph (1 M nodes, 20 M edges), batch 1024, fanout [15,10,5].

  python benchmarks/bench_cpu_sampler.py ours | reference
"""
import json, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
impl=sys.argv[1]
N, E, B, fan, iters = 1_000_000, 20_000_000, 1024, [15,10,5], 30
g=torch.Generator().manual_seed(0)
src=torch.randint(0,N,(E//2,),generator=g); dst=(src+torch.randint(1,5000,(E//2,),generator=g))%N
ei=torch.stack([torch.cat([src,dst]),torch.cat([dst,src])])
seeds=[torch.randint(0,N,(B,),generator=g) for _ in range(iters+3)]
if impl=='reference':
  sys.path.insert(0, os.path.join(ROOT,'baseline','shims')); sys.path.insert(0, os.path.join(ROOT,'baseline','_ref'))
  import graphlearn_torch as glt
  topo=glt.data.Topology(ei, input_layout='COO')
  graph=glt.data.Graph(topo,'CPU')
  s=glt.sampler.NeighborSampler(graph, fan, device=torch.device('cpu'))
else:
  import graphlearn_for_pytorch_b200 as glt
  topo=glt.data.Topology(ei, layout='CSR', num_nodes=N)
  graph=glt.data.Graph(topo,'CPU')
  s=glt.sampler.NeighborSampler(graph, fan, device=torch.device('cpu'), seed=1)
for sd in seeds[:3]: s.sample_from_nodes(sd)
t=time.time(); edges=0
for sd in seeds[3:]:
  edges+=s.sample_from_nodes(sd).row.numel()
dt=time.time()-t
print(json.dumps({'impl':impl,'cpu_threads':torch.get_num_threads(),'M_edges_per_s':edges/dt/1e6,'ms_per_batch':dt/iters*1e3,'edges_per_batch':edges/iters}))
