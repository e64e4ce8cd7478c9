"""Unsupervised GraphSAGE with link sampling + binary negatives
(counterpart of the reference's examples/graph_sage_unsup_ppi.py)."""
import torch
import torch.nn.functional as F

from common import glt, synthetic_homo
from graphlearn_for_pytorch_b200.models import GraphSAGE
from graphlearn_for_pytorch_b200.sampler import NegativeSampling

cuda = torch.cuda.is_available()
device = torch.device('cuda', 0) if cuda else torch.device('cpu')
ei, x, _ = synthetic_homo(20_000, 200_000, feat_dim=50, num_classes=4)
ds = glt.data.Dataset()
ds.init_graph(ei, graph_mode='CUDA' if cuda else 'CPU')
ds.init_node_features(x, with_gpu=cuda, split_ratio=1.0 if cuda else 0.0)
loader = glt.loader.LinkNeighborLoader(ds, [10, 10], edge_label_index=ei[:, :20_000], batch_size=512, shuffle=True,
                                       neg_sampling=NegativeSampling('binary', 1), device=device)
model = GraphSAGE(50, 64, 64, num_layers=2).to(device)
opt = torch.optim.Adam(model.parameters(), lr=5e-3)
for epoch in range(2):
  tot = 0.0
  for b in loader:
    h = model(b.x, b.edge_index)
    src, dst = b.edge_label_index[1], b.edge_label_index[0]
    logit = (h[src] * h[dst]).sum(-1)
    loss = F.binary_cross_entropy_with_logits(logit, b.edge_label.float().to(device))
    opt.zero_grad(); loss.backward(); opt.step()
    tot += float(loss.detach())
  print(f'epoch {epoch}: loss {tot / len(loader):.4f}')
