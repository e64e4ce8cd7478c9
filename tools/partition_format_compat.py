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

# ---- the other direction: partitions written by the REFERENCE's partitioner (with edge features) are read by this
# library's load_partition / DistDataset.load -- an existing partitioned dataset can be used as it is
N,E=2000,20000
g=torch.Generator().manual_seed(0)
ei=torch.randint(0,N,(2,E),generator=g)
x=torch.randn(N,16,generator=g)
ef=torch.randn(E,4,generator=g)
d=tempfile.mkdtemp()
p=rglt.partition.RandomPartitioner(d, 2, N, ei, node_feat=x, edge_feat=ef, edge_assign_strategy='by_src', chunk_size=1000)
p.partition()
print(sorted(os.listdir(d)), sorted(os.listdir(os.path.join(d,'part0'))))
r = rglt.partition.load_partition(d, 1)
o = glt.partition.load_partition(d, 1)
re = torch.stack(list(r[2].edge_index)); oe = torch.stack(list(o[2].edge_index)) if isinstance(o[2].edge_index,(tuple,list)) else o[2].edge_index
print('same edges', torch.equal(oe, re), 'eids', torch.equal(o[2].eids, r[2].eids), 'nfeat', torch.equal(o[3].feats, r[3].feats), torch.equal(o[3].ids, r[3].ids),
      'efeat', torch.equal(o[4].feats, r[4].feats), torch.equal(o[4].ids, r[4].ids),
      'pbs', torch.equal(torch.as_tensor(o[5]), torch.as_tensor(r[5])), torch.equal(torch.as_tensor(o[6]), torch.as_tensor(r[6])))
# and our DistDataset on the reference-written directory
ds = glt.distributed.DistDataset()
ds.load(d, 1, graph_mode='CPU')
print('DistDataset ok', ds.num_partitions, ds.partition_idx, ds.node_features is not None)

print('FORMAT OK')
