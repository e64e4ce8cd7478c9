"""Smoke-run the CPU-capable example scripts (tiny settings) so they cannot rot."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=300):
  env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get('CUDA_VISIBLE_DEVICES', ''))
  out = subprocess.run([sys.executable] + args, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)
  assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
  return out.stdout


def test_llm_on_subgraphs_stub_backend():
  out = _run(['examples/gpt/arxiv_llm.py', '--backend', 'stub', '--batches', '8'])
  assert 'link-prediction accuracy' in out
