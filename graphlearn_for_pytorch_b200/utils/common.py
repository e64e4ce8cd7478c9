"""Misc helpers: seeding, ports, dict merging, tensor-file appends, checkpoints.

Parity: reference python/utils/common.py:30-234.
"""
import os
import pickle
import random
import socket
from typing import Any, Dict, List, Optional

import numpy as np
import torch


def ensure_dir(dir_path: str):
  os.makedirs(dir_path, exist_ok=True)


def seed_everything(seed: int):
  random.seed(seed)
  np.random.seed(seed % (2 ** 32))
  torch.manual_seed(seed)
  if torch.cuda.is_available():
    torch.cuda.manual_seed_all(seed)


def get_free_port(host: str = '127.0.0.1') -> int:
  s = socket.socket()
  s.bind((host, 0))
  port = s.getsockname()[1]
  s.close()
  return port


def get_free_port_block(n: int = 8, host: str = '127.0.0.1', tries: int = 64) -> int:
  """First port of `n` CONSECUTIVE free ports (callers that derive `port + k` rendezvous points from one base port
  -- sampling worker groups, per-client channels -- would otherwise collide with unrelated sockets now and then)."""
  import random
  rng = random.Random()
  for _ in range(tries):
    base = rng.randrange(20000, 60000 - n)
    socks = []
    try:
      for k in range(n):
        sk = socket.socket()
        sk.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 0)
        sk.bind((host, base + k))
        socks.append(sk)
      return base
    except OSError:
      continue
    finally:
      for sk in socks:
        sk.close()
  return get_free_port(host)


def merge_dict(in_dict: Dict[Any, Any], out_dict: Dict[Any, List[Any]]):
  for k, v in in_dict.items():
    out_dict.setdefault(k, []).append(v)


def count_dict(in_dict: Dict[Any, Any], out_dict: Dict[Any, List[int]], target_len: int):
  """Append len(v) per key, padding keys unseen in earlier hops with zeros."""
  for k, v in in_dict.items():
    vals = out_dict.setdefault(k, [])
    vals.extend([0] * (target_len - len(vals) - 1))
    vals.append(len(v))
  for k, vals in out_dict.items():
    vals.extend([0] * (target_len - len(vals)))


def index_select(data, index):
  if data is None:
    return None
  if isinstance(data, dict):
    return {k: index_select(v, index) for k, v in data.items()}
  if isinstance(data, (list, tuple)):
    return type(data)(index_select(v, index) for v in data)
  if isinstance(index, tuple):
    start, end = index
    return data[start:end]
  return data[index]


def merge_hetero_sampler_output(in_sample, out_sample, device, edge_dir: str = 'out'):
  """Merge `in_sample` into `out_sample` (both HeteroSamplerOutput) with node dedup per type."""
  def _unique_with_inverse(a, b):
    cat = torch.cat([a, b])
    uniq, inv = torch.unique(cat, return_inverse=True)
    return cat, uniq, inv
  for ntype, nodes in in_sample.node.items():
    if ntype not in out_sample.node:
      out_sample.node[ntype] = nodes
    else:
      out_sample.node[ntype] = torch.unique(torch.cat([out_sample.node[ntype], nodes]))
  for etype, rows in in_sample.row.items():
    cols = in_sample.col[etype]
    if etype in out_sample.row:
      out_sample.row[etype] = torch.cat([out_sample.row[etype], rows])
      out_sample.col[etype] = torch.cat([out_sample.col[etype], cols])
    else:
      out_sample.row[etype] = rows
      out_sample.col[etype] = cols
    if in_sample.edge is not None and etype in in_sample.edge:
      if out_sample.edge is None:
        out_sample.edge = {}
      out_sample.edge[etype] = torch.cat([out_sample.edge[etype], in_sample.edge[etype]]) \
        if etype in out_sample.edge else in_sample.edge[etype]
  return out_sample


def format_hetero_sampler_output(in_sample, edge_dir: str = 'out'):
  """Make sure every edge type's endpoint node types exist in the node dict."""
  for k in list(in_sample.row.keys()):
    for t in (k[0], k[-1]):
      if t not in in_sample.node:
        in_sample.node[t] = torch.empty(0, dtype=torch.int64, device=in_sample.row[k].device)
  return in_sample


# ---- append-only tensor files (on-disk partition format, chunked features) ----
def append_tensor_to_file(filename: str, tensor: torch.Tensor):
  with open(filename, 'ab') as f:
    pickle.dump(tensor.cpu(), f, pickle.HIGHEST_PROTOCOL)


def load_and_concatenate_tensors(filename: str, device=None) -> Optional[torch.Tensor]:
  chunks = []
  with open(filename, 'rb') as f:
    while True:
      try:
        chunks.append(pickle.load(f))
      except EOFError:
        break
  if not chunks:
    return None
  out = torch.cat(chunks, dim=0)
  return out.to(device) if device is not None else out


def default_id_select(srcs: torch.Tensor, p_mask: torch.Tensor, node_pb=None) -> torch.Tensor:
  return torch.masked_select(srcs, p_mask)


def default_id_filter(node_pb: torch.Tensor, partition_idx: int) -> torch.Tensor:
  return torch.where(node_pb == partition_idx)[0]


# ---- checkpoints (reference utils/common.py:177-234) ----
def save_ckpt(ckpt_seq: int, ckpt_dir: str, model, optimizer=None, epoch: float = 0,
              extra: Optional[Dict[str, Any]] = None):
  """Write model_seq_{n}.ckpt with model/optimizer state (+ loader/RNG state in `extra`)."""
  ensure_dir(ckpt_dir)
  path = os.path.join(ckpt_dir, f'model_seq_{ckpt_seq}.ckpt')
  state = {
    'seq': ckpt_seq,
    'epoch': epoch,
    'model_state_dict': model.state_dict(),
    'optimizer_state_dict': optimizer.state_dict() if optimizer is not None else None,
  }
  if extra:
    state['extra'] = extra
  torch.save(state, path)
  return path


def load_ckpt(ckpt_seq: int, ckpt_dir: str, model, optimizer=None, return_extra: bool = False):
  """Restore a checkpoint; returns the stored epoch, or -1 when it does not exist."""
  path = os.path.join(ckpt_dir, f'model_seq_{ckpt_seq}.ckpt')
  if not os.path.isfile(path):
    return (-1, None) if return_extra else -1
  state = torch.load(path, map_location='cpu', weights_only=False)
  model.load_state_dict(state['model_state_dict'])
  if optimizer is not None and state.get('optimizer_state_dict') is not None:
    optimizer.load_state_dict(state['optimizer_state_dict'])
  if return_extra:
    return state.get('epoch', -1), state.get('extra')
  return state.get('epoch', -1)


class RandomSeedManager(object):
  """Process-wide optional seed for samplers created without an explicit `seed`
  (parity: reference include/common.h:36-65).  With a seed set, every sampler draws from
  Philox streams derived from it, so whole runs are reproducible."""
  _seed = None
  _count = 0

  @classmethod
  def set_seed(cls, seed: int):
    cls._seed = int(seed)
    cls._count = 0

  @classmethod
  def get_seed(cls):
    return cls._seed

  @classmethod
  def next_seed(cls):
    """A fresh per-object seed: deterministic when a global seed is set, random otherwise."""
    if cls._seed is None:
      return int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    cls._count += 1
    return (cls._seed * 1000003 + cls._count * 7919) % (2 ** 31 - 1)
