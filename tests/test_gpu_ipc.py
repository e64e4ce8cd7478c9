"""Cross-process sharing of GPU-resident data: runs tests/mp/feature_ipc_check.py (a parent that
builds Feature / UnifiedTensor / Graph and two spawned readers) as a script, like the other
multi-process GPU checks."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_feature_graph_ipc_to_spawned_readers():
  out = subprocess.run([sys.executable, os.path.join(HERE, 'mp', 'feature_ipc_check.py')],
                       capture_output=True, text=True, timeout=240)
  assert out.returncode == 0 and 'IPC_OK' in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
