"""Multi-GPU checks wrapped for `pytest -m gpu`: they run whenever the lease has >= 2 GPUs and skip on a
single-GPU box, so any multi-GPU driver lease exercises the NVLink data plane (peer gather, peer sampling ==
full-graph sampling, fused layer-1 numerics across shards, engine gradient sync, P2P DistNeighborLoader)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _ngpu():
  import torch
  return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
def test_p2p_data_plane_two_ranks():
  if _ngpu() < 2:
    pytest.skip('needs >= 2 GPUs')
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
         '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
         os.path.join(ROOT, 'tests', 'mp', 'p2p_check.py')]
  r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
  assert r.returncode == 0, r.stdout[-3000:] + '\n' + r.stderr[-3000:]
  assert 'ALL OK' in r.stdout or 'all ok' in r.stdout.lower(), r.stdout[-2000:]


@pytest.mark.gpu
def test_feature_ipc_across_two_gpus():
  if _ngpu() < 2:
    pytest.skip('needs >= 2 GPUs')
  r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'mp', 'feature_ipc_check.py')], cwd=ROOT,
                     capture_output=True, text=True, timeout=600)
  assert r.returncode == 0, r.stdout[-3000:] + '\n' + r.stderr[-3000:]
