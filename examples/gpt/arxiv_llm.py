"""LLM-over-sampled-subgraphs (counterpart of the reference's examples/gpt/arxiv.py + utils.py).

The framework part is identical to the reference example: a `LinkNeighborLoader` with binary
negative sampling draws a 2-hop subgraph around two positive and two negative candidate edges,
the subgraph is verbalised into a prompt (paper titles + edge list + 4 yes/no questions) and the
prompt goes to a language model.  What differs is that the model is pluggable, so the example also
runs where no LLM endpoint is reachable:

  --backend openai   OpenAI-compatible chat endpoint (needs the `openai` package and OPENAI_API_KEY)
  --backend hf       local `transformers` text-generation pipeline (--model path/to/weights)
  --backend stub     offline structural baseline: answers "yes" when the two endpoints share a
                     neighbour in the sampled subgraph (lets you check prompts + scoring end to end)

  python examples/gpt/arxiv_llm.py --backend stub --batches 50
With real data: point --root at a directory holding titles.csv.gz / ids.csv.gz / edges.csv.gz
(the arxiv_2023 layout used by the reference); without it a synthetic citation graph is used.
"""
import argparse
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import glt  # noqa: E402


def link_prediction_prompt(batch, titles, reason=False):
  n = batch.node.numel()
  edges = batch.edge_index.t().unique(dim=0).tolist()
  lines = [f'This is a directed subgraph of an arxiv citation network with {n} nodes numbered from 0 to {n - 1}.',
           'The titles of each paper:']
  lines += [f"node {i} is '{titles[i]}'" for i in range(n)]
  lines.append(f'The sampled subgraph of the network is {edges} where the first number indicates the source '
               'node and the second the destination node.')
  lines.append('Hint: the direction of an edge carries information about temporality.')
  lines.append('According to the principles of citation-network construction and the given subgraph, answer:')
  cand = batch.edge_label_index.t().tolist()
  order = torch.randperm(len(cand)).tolist()           # do not leak the label through the position
  for q, j in enumerate(order, 1):
    lines.append(f'Question {q}: predict whether there tends to form an edge {cand[j]}.')
  lines.append('Answer yes or no for every question' + (' and show your reasoning.' if reason
                                                        else " and don't show any reasoning process."))
  return '\n'.join(lines), [cand[j] for j in order], [float(batch.edge_label[j]) for j in order]


def node_classification_prompt(batch):
  n = batch.node.numel()
  lines = [f'This is a directed subgraph of an arxiv citation network with {n} nodes numbered from 0 to {n - 1}.',
           f'The subgraph has {batch.edge_index.shape[1]} edges.']
  for i in range(1, n):
    feat = ','.join(f'{v:.3f}' for v in batch.x[i].tolist())
    lines.append(f'The feature of node {i} is [{feat}] and the node label is {int(batch.y[i])}.')
  lines.append(f'The edges of the subgraph are {batch.edge_index.t().tolist()} (source, destination).')
  feat0 = ','.join(f'{v:.3f}' for v in batch.x[0].tolist())
  lines.append(f"Question: predict the label for node 0, whose feature is [{feat0}]. Give the label only.")
  return '\n'.join(lines)


class StubLLM(object):
  """Deterministic structural baseline that speaks the same protocol (prompt in, text out)."""

  def __call__(self, prompt: str) -> str:
    m = re.search(r'sampled subgraph of the network is (\[.*?\]) where', prompt, re.S)
    edges = eval(m.group(1)) if m else []          # noqa: S307 (our own serialisation)
    nbrs = {}
    for s, d in edges:
      nbrs.setdefault(s, set()).add(d)
      nbrs.setdefault(d, set()).add(s)
    out = []
    for q, (a, b) in enumerate(re.findall(r'form an edge \[(\d+), (\d+)\]', prompt), 1):
      a, b = int(a), int(b)
      common = (nbrs.get(a, set()) & nbrs.get(b, set())) - {a, b}
      out.append(f"Question {q}: {'yes' if common or b in nbrs.get(a, set()) else 'no'}")
    return '\n'.join(out)


def make_llm(args):
  if args.backend == 'stub':
    return StubLLM()
  if args.backend == 'openai':
    from openai import OpenAI
    client = OpenAI()

    def call(prompt):
      r = client.chat.completions.create(messages=[{'role': 'user', 'content': prompt}], model=args.model)
      return r.choices[0].message.content
    return call
  from transformers import pipeline
  pipe = pipeline('text-generation', model=args.model, device=0 if torch.cuda.is_available() else -1)
  return lambda prompt: pipe(prompt, max_new_tokens=64, return_full_text=False)[0]['generated_text']


def parse_answers(text: str, n: int):
  ans = re.findall(r'\b(yes|no)\b', text.lower())
  return [1.0 if a == 'yes' else 0.0 for a in ans[:n]] + [0.0] * max(0, n - len(ans))


def load_data(root):
  if root and os.path.exists(os.path.join(root, 'edges.csv.gz')):
    import pandas as pd
    titles = [t[0] for t in pd.read_csv(os.path.join(root, 'titles.csv.gz')).to_numpy()]
    edge_index = torch.from_numpy(pd.read_csv(os.path.join(root, 'edges.csv.gz')).to_numpy()).t().contiguous()
    return edge_index, titles
  # synthetic "citation" graph: papers cite earlier papers of their own topic
  g = torch.Generator().manual_seed(0)
  n, topics = 5000, 40
  topic = torch.randint(0, topics, (n,), generator=g)
  src = torch.arange(200, n).repeat_interleave(6)
  cand = (torch.rand(src.numel(), generator=g) * src.float()).long()
  same = topic[cand] == topic[src]
  src, cand = src[same | (torch.rand(src.numel(), generator=g) < 0.15)], cand[same | (torch.rand(src.numel(), generator=g) < 0.15)]
  k = min(src.numel(), cand.numel())
  titles = [f'Paper {i} on topic {int(topic[i])}' for i in range(n)]
  return torch.stack([src[:k], cand[:k]]), titles


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--backend', default='stub', choices=['stub', 'openai', 'hf'])
  ap.add_argument('--model', default='gpt-4-1106-preview')
  ap.add_argument('--root', default='')
  ap.add_argument('--batches', type=int, default=20)
  ap.add_argument('--reason', action='store_true')
  args = ap.parse_args()
  edge_index, titles = load_data(args.root)
  n = int(edge_index.max()) + 1
  ds = glt.data.Dataset()
  ds.init_graph(edge_index=edge_index, graph_mode='CPU', directed=True, num_nodes=n)
  ds.init_node_features(torch.arange(n, dtype=torch.float32).unsqueeze(1), sort_func=glt.data.sort_by_in_degree,
                        split_ratio=0.0, with_gpu=False)
  loader = glt.loader.LinkNeighborLoader(ds, [12, 6], neg_sampling=glt.sampler.NegativeSampling('binary'),
                                         batch_size=2, drop_last=True, shuffle=True, device=torch.device('cpu'))
  llm = make_llm(args)
  hit = tot = 0
  for i, batch in enumerate(loader):
    if i >= args.batches:
      break
    if batch.edge_index.shape[1] < 5:
      continue
    names = [titles[int(v)] for v in batch.node]
    prompt, cand, truth = link_prediction_prompt(batch, names, args.reason)
    pred = parse_answers(llm(prompt), len(cand))
    hit += sum(int(p == t) for p, t in zip(pred, truth))
    tot += len(cand)
    if i == 0:
      print(prompt[:1200] + ('...' if len(prompt) > 1200 else ''))
  print(f'[{args.backend}] link-prediction accuracy over {tot} candidate edges: {hit / max(tot, 1):.3f}')


if __name__ == '__main__':
  main()
