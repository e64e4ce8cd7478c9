"""CPU-mode heterogeneous NeighborSampler (graphs in host memory, device=cpu): this library vs the unmodified reference
(baseline/_ref) on the same synthetic IGBH-like schema (paper / author / institute, 5 relations, 400 k papers),
paper seeds, batch 1024, fan-out [15,10,5].

  python benchmarks/bench_cpu_hetero_sampler.py ours | reference
"""
import json
import os
import sys
import time

import torch

impl = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if impl=='reference':
  sys.path.insert(0, os.path.join(ROOT,'baseline','shims')); sys.path.insert(0, os.path.join(ROOT,'baseline','_ref'))
  import graphlearn_torch as glt
else:
  sys.path.insert(0, ROOT); import graphlearn_for_pytorch_b200 as glt
g=torch.Generator().manual_seed(0)
NP,NA,NI=400_000,800_000,20_000
def e(ns,nd,m): return torch.stack([torch.randint(0,ns,(m,),generator=g), torch.randint(0,nd,(m,),generator=g)])
edges={('paper','cites','paper'):e(NP,NP,4_000_000), ('paper','written_by','author'):e(NP,NA,3_000_000),
       ('author','rev_written_by','paper'):None, ('author','affiliated_to','institute'):e(NA,NI,800_000),
       ('institute','rev_affiliated_to','author'):None}
edges[('author','rev_written_by','paper')]=edges[('paper','written_by','author')].flip(0)
edges[('institute','rev_affiliated_to','author')]=edges[('author','affiliated_to','institute')].flip(0)
graphs={}
for et,ei in edges.items():
  if impl=='reference':
    topo=glt.data.Topology(ei, input_layout='COO')
  else:
    topo=glt.data.Topology(ei, layout='CSR')
  graphs[et]=glt.data.Graph(topo,'CPU')
s=glt.sampler.NeighborSampler(graphs,[15,10,5],device=torch.device('cpu'))
seeds=[torch.randint(0,NP,(1024,),generator=g) for _ in range(13)]
from_nodes = (lambda sd: s.sample_from_nodes(glt.sampler.NodeSamplerInput(node=sd, input_type='paper')))
for sd in seeds[:3]: from_nodes(sd)
t=time.time(); ne=0
for sd in seeds[3:]:
  o=from_nodes(sd); ne+=sum(v.numel() for v in o.row.values())
dt=time.time()-t
print(json.dumps({'impl':impl,'ms_per_batch':dt/10*1e3,'M_edges_per_s':ne/dt/1e6,'edges_per_batch':ne/10}))
