"""SEAL link prediction: enclosing-subgraph sampling + DRNL + DGCNN
(counterpart of the reference's examples/seal_link_pred.py)."""
import argparse

import torch
import torch.nn.functional as F

from common import glt, synthetic_homo
from graphlearn_for_pytorch_b200.models import DGCNN, drnl_node_labeling
from graphlearn_for_pytorch_b200.sampler import NeighborSampler, NodeSamplerInput, RandomNegativeSampler

ap = argparse.ArgumentParser()
ap.add_argument('--links', type=int, default=2000, help='positive (and negative) links')
ap.add_argument('--epochs', type=int, default=2)
args = ap.parse_args()
L = args.links
cuda = torch.cuda.is_available()
device = torch.device('cuda', 0) if cuda else torch.device('cpu')
ei, _, _ = synthetic_homo(3_000, 24_000, feat_dim=4, num_classes=2)
topo = glt.data.Topology(ei, layout='CSR', num_nodes=3000)
graph = glt.data.Graph(topo, 'CUDA' if cuda else 'CPU')
sampler = NeighborSampler(graph, [-1], device=device)          # 1-hop enclosing subgraphs
neg = RandomNegativeSampler(graph, 'CUDA' if cuda else 'CPU').sample(L, padding=True).cpu()
pos = ei[:, torch.randperm(ei.shape[1])[:L]]
links = torch.cat([pos, neg], 1)
labels = torch.cat([torch.ones(L), torch.zeros(L)])
model = DGCNN(num_labels=200, hidden=32, num_layers=3, k=30).to(device)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)


def extract(batch_links):
  zs, eis, batch, off = [], [], [], 0
  for g, (s, d) in enumerate(batch_links.t().tolist()):
    out = sampler.subgraph(NodeSamplerInput(torch.tensor([s, d])))
    sub_ei = torch.stack([out.row, out.col])
    m = out.metadata
    keep = ~(((sub_ei[0] == m[0]) & (sub_ei[1] == m[1])) | ((sub_ei[0] == m[1]) & (sub_ei[1] == m[0])))
    sub_ei = sub_ei[:, keep]                                    # hide the target link
    n = out.node.numel()
    zs.append(drnl_node_labeling(sub_ei, int(m[0]), int(m[1]), n, max_z=199))
    eis.append(sub_ei + off)
    batch.append(torch.full((n,), g, device=device))
    off += n
  return torch.cat(zs), torch.cat(eis, 1), torch.cat(batch)


for epoch in range(args.epochs):
  perm = torch.randperm(links.shape[1])
  tot, correct = 0.0, 0
  for i in range(0, perm.numel(), 32):
    idx = perm[i:i + 32]
    z, sub_ei, batch = extract(links[:, idx])
    logit = model(z.to(device), sub_ei.to(device), batch, idx.numel())
    y = labels[idx].to(device)
    loss = F.binary_cross_entropy_with_logits(logit, y)
    opt.zero_grad(); loss.backward(); opt.step()
    tot += float(loss.detach()) * idx.numel(); correct += int(((logit > 0).float() == y).sum())
  print(f'epoch {epoch}: loss {tot / perm.numel():.4f} acc {correct / perm.numel():.4f}')
