"""On-disk partition format check: partitions written by THIS library's RandomPartitioner are read by the unmodified
reference's `load_partition` (baseline/_ref) and by ours, and both see the same tensors.

  python tools/partition_format_compat.py      # prints 'FORMAT OK' when every comparison holds
"""
import os, sys, torch, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graphlearn_for_pytorch_b200 as glt
d=tempfile.mkdtemp()
N,E=2000,20000
g=torch.Generator().manual_seed(0)
ei=torch.randint(0,N,(2,E),generator=g)
x=torch.randn(N,16,generator=g)
p=glt.partition.RandomPartitioner(d, 2, N, ei, node_feat=x, edge_feat=None, edge_assign_strategy='by_src', chunk_size=1000)
p.partition()
print(sorted(os.listdir(d)), sorted(os.listdir(os.path.join(d,'part0'))))
# now load with the reference
sys.path.insert(0, os.path.join(ROOT,'baseline','shims')); sys.path.insert(0, os.path.join(ROOT,'baseline','_ref'))
import graphlearn_torch as rglt
out = rglt.partition.load_partition(d, 0)
print(type(out), len(out))
num_parts, idx, graph, nfeat, efeat, npb, epb = out
o = glt.partition.load_partition(d, 0)
re = torch.stack(list(graph.edge_index)) if isinstance(graph.edge_index, (tuple, list)) else graph.edge_index
print('reference loaded: parts', num_parts, 'edges', tuple(re.shape), 'feats', tuple(nfeat.feats.shape), 'pb', tuple(npb.shape))
print('same edges', torch.equal(torch.stack(list(o[2].edge_index)) if isinstance(o[2].edge_index,(tuple,list)) else o[2].edge_index, re), 'same eids', torch.equal(o[2].eids, graph.eids),
      'same feats', torch.equal(o[3].feats, nfeat.feats), 'same ids', torch.equal(o[3].ids, nfeat.ids),
      'same node_pb', torch.equal(torch.as_tensor(o[5]), torch.as_tensor(npb)), 'same edge_pb', torch.equal(torch.as_tensor(o[6]), torch.as_tensor(epb)))

assert out is not None
print('FORMAT OK')
