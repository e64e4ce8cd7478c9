"""Unsupervised bipartite GraphSAGE (user-item) with hetero link sampling and triplet negatives --
counterpart of the reference's examples/hetero/bipartite_sage_unsup.py."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import glt  # noqa: E402
from graphlearn_for_pytorch_b200.models import RGNN  # noqa: E402
from graphlearn_for_pytorch_b200.sampler import NegativeSampling  # noqa: E402

cuda = torch.cuda.is_available()
device = torch.device('cuda', 0) if cuda else torch.device('cpu')
g = torch.Generator().manual_seed(0)
nu, ni = 2000, 1000
u2i = torch.stack([torch.randint(0, nu, (20000,), generator=g), torch.randint(0, ni, (20000,), generator=g)])
ds = glt.data.Dataset(edge_dir='out')
ds.init_graph({('user', 'u2i', 'item'): u2i, ('item', 'rev_u2i', 'user'): u2i.flip(0)},
              graph_mode='CUDA' if cuda else 'CPU', num_nodes={'user': nu, 'item': ni})
ds.init_node_features({'user': torch.randn(nu, 32, generator=g), 'item': torch.randn(ni, 32, generator=g)},
                      with_gpu=cuda, split_ratio=1.0 if cuda else 0.0)
et = ('user', 'u2i', 'item')
loader = glt.loader.LinkNeighborLoader(ds, [8, 4], edge_label_index=(et, u2i[:, :8000]), batch_size=256, shuffle=True,
                                       neg_sampling=NegativeSampling('triplet', 1), device=device)
first = next(iter(loader))
user_enc = RGNN(list(first.edge_index_dict.keys()), 32, 64, 64, num_layers=2, node_type='user').to(device)
item_enc = RGNN(list(first.edge_index_dict.keys()), 32, 64, 64, num_layers=2, node_type='item').to(device)
opt = torch.optim.Adam(list(user_enc.parameters()) + list(item_enc.parameters()), lr=5e-3)
for epoch in range(2):
  tot = 0.0
  for b in loader:
    hu = user_enc(b.x_dict, b.edge_index_dict)
    hi = item_enc(b.x_dict, b.edge_index_dict)
    src, pos, neg = b['user'].src_index, b['item'].dst_pos_index, b['item'].dst_neg_index[:, 0]
    loss = F.softplus(-(hu[src] * hi[pos]).sum(-1)).mean() + F.softplus((hu[src] * hi[neg]).sum(-1)).mean()
    opt.zero_grad(); loss.backward(); opt.step()
    tot += float(loss.detach())
  print(f'epoch {epoch}: loss {tot / len(loader):.4f}')
