"""`GraphSageTrainer`: the device-agnostic counterpart of `GraphSageEngine`.

Same step-level API (`train_step(seeds)`, `evaluate_batch(seeds)`, `state_dict()` / `load_state_dict()`), built
from the public pieces -- `NeighborSampler`, `Feature`, the eager `GraphSAGE` module and `torch.optim.Adam` -- so it
runs on a CPU as well as on any GPU.  Two uses:

* code that must run without a B200 (CI, notebooks) keeps the engine's call pattern and switches class;
* it is the *semantic reference* of the fused engine: same sampler stream, same model family, autograd backward.

(The reference has no engine; its examples hand-write this loop, e.g. examples/train_sage_ogbn_products.py:61-110.)
"""
from typing import List, Optional

import torch
import torch.nn.functional as F

from ..sampler import NeighborSampler, NodeSamplerInput
from .sage import GraphSAGE


class GraphSageTrainer(object):
  def __init__(self, graph, features, labels: torch.Tensor, in_dim: int, fanouts: List[int] = (15, 10, 5),
               hidden: int = 256, num_classes: int = 47, lr: float = 3e-3, weight_decay: float = 0.0, seed: int = 0,
               device: Optional[torch.device] = None):
    self.device = torch.device(device) if device is not None else torch.device('cpu')
    self.features, self.labels = features, labels.to(self.device)
    self.fanouts = list(fanouts)
    self.sampler = NeighborSampler(graph, self.fanouts, device=self.device, seed=seed)
    torch.manual_seed(seed)
    self.model = GraphSAGE(in_dim, hidden, num_classes, len(self.fanouts)).to(self.device)
    self.opt = torch.optim.Adam(self.model.parameters(), lr=lr, weight_decay=weight_decay)
    self.step_idx = 0
    self.loss = torch.zeros((), device=self.device)

  def _batch(self, seeds: torch.Tensor):
    out = self.sampler.sample_from_nodes(NodeSamplerInput(node=seeds.to(self.sampler.device)))
    x = self.features[out.node] if not isinstance(self.features, torch.Tensor) else self.features[out.node.cpu()]
    x = x.to(self.device, dtype=torch.float32)
    edge_index = torch.stack([out.row, out.col]).to(self.device)
    n_seed = out.batch.numel() if out.batch is not None else seeds.numel()
    y = self.labels[out.node[:n_seed].to(self.device)]
    return x, edge_index, out.num_sampled_nodes, out.num_sampled_edges, y, n_seed

  def train_step(self, seeds: torch.Tensor) -> torch.Tensor:
    """One optimisation step on a batch of seed ids; returns the (device) mean NLL loss."""
    self.model.train()
    x, ei, nn_, ne_, y, n = self._batch(seeds)
    logits = self.model(x, ei, nn_, ne_)[:n]
    loss = F.cross_entropy(logits, y)
    self.opt.zero_grad(set_to_none=True)
    loss.backward()
    self.opt.step()
    self.step_idx += 1
    self.loss = loss.detach()
    return self.loss

  @torch.no_grad()
  def evaluate_batch(self, seeds: torch.Tensor):
    """(loss, #correct, #seeds) without updating parameters -- same tuple as GraphSageEngine.evaluate_batch."""
    self.model.eval()
    x, ei, nn_, ne_, y, n = self._batch(seeds)
    logits = self.model(x, ei, nn_, ne_)[:n]
    return float(F.cross_entropy(logits, y)), int((logits.argmax(1) == y).sum()), int(n)

  def flush(self):          # API compatibility with the pipelined engine
    return None

  def close(self):
    return None

  def state_dict(self):
    s = {'model': self.model.state_dict(), 'optimizer': self.opt.state_dict(), 'step': self.step_idx}
    if hasattr(self.sampler, 'state_dict'):
      s['sampler'] = self.sampler.state_dict()
    return s

  def load_state_dict(self, s):
    self.model.load_state_dict(s['model'])
    self.opt.load_state_dict(s['optimizer'])
    self.step_idx = int(s.get('step', 0))
    if 'sampler' in s and hasattr(self.sampler, 'load_state_dict'):
      self.sampler.load_state_dict(s['sampler'])
