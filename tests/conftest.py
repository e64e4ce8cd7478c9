import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: test needs a real CUDA device (B200)')


@pytest.fixture(scope='session')
def glt():
  import graphlearn_for_pytorch_b200 as g
  assert g.ops.has_native(), f'native extension failed to load: {g.ops._load_error!r}'
  return g


@pytest.fixture(scope='session')
def native(glt):
  return glt.ops.require_native()
