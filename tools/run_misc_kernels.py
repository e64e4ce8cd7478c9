"""Launch every non-engine kernel of the library once or twice on a products-shape graph so that one
`ncu --set full -k regex:...` capture covers them (evidence pack: gather, one-hop / negative / subgraph / random-walk
/ sample_prob kernels, hetero grouped sampling, MXFP8 gather).  Not a benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphlearn_for_pytorch_b200 as glt  # noqa: E402
from graphlearn_for_pytorch_b200.sampler import NeighborSampler, NodeSamplerInput, RandomNegativeSampler  # noqa: E402
from graphlearn_for_pytorch_b200.utils.synthetic import rmat_edges  # noqa: E402

dev = torch.device('cuda', 0)
N, E = 2_449_029, 123_718_280
ei = rmat_edges(N, E // 2, seed=0, device=dev)
ei = torch.cat([ei, ei.flip(0)], 1)
topo = glt.data.Topology(ei, layout='CSR', num_nodes=N)
del ei
g = glt.data.Graph(topo, 'CUDA', 0)
feats = torch.randn(N, 128, device=dev).to(torch.bfloat16)
ut = glt.data.UnifiedTensor(0, torch.bfloat16)
ut.append_shared_tensor(feats)
q = glt.data.quantize_mxfp8(feats[:1 << 20])
utq = glt.data.UnifiedTensor(0, torch.uint8)
utq.append_shared_tensor(q)
s = NeighborSampler(g, [15, 10, 5], device=dev, seed=1, with_edge=False)
neg = RandomNegativeSampler(g, 'CUDA', seed=2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for it in range(2):
  ids = torch.randint(0, N, (400_000,), device=dev)
  ut[ids]                                                           # k_gather_vec
  utq._table().gather_mxfp8(torch.randint(0, 1 << 20, (400_000,), device=dev), 128)   # k_gather_mxfp8
  seeds = torch.randint(0, N, (1024,), device=dev)
  s.sample_one_hop(seeds, 15)                                       # k_sample_one_hop
  s.sample_from_nodes(NodeSamplerInput(seeds))                      # arena kernels + k_ell_to_coo
  neg.sample(1 << 20, trials_num=5)                                 # k_negative_sample
  s.subgraph(NodeSamplerInput(torch.randint(0, N, (512,), device=dev)))   # k_subgraph_count / fill
  s.random_walk(torch.randint(0, N, (100_000,), device=dev), 10) if hasattr(s, 'random_walk') else None
  try:
    s.sample_prob(NodeSamplerInput(seeds), N)                       # k_nbr_prob
  except Exception:
    pass
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
