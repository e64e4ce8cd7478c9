"""Native C++ unit tests (tests/cpp/*.cc) compiled and run without Python in the loop, plus a
ThreadSanitizer build of the shm ring's producer/consumer test (the reference has no
sanitizer job)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which('g++') is None, reason='no host compiler')
def test_native_cpp_suites_and_tsan():
  out = subprocess.run(['bash', os.path.join(ROOT, 'scripts', 'run_cpp_ut.sh'), 'tsan'], capture_output=True,
                       text=True, timeout=900)
  assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
  assert out.stdout.count('SHM_QUEUE_OK') == 2 and 'ThreadSanitizer' not in out.stderr
  assert 'CPU_OPS_OK' in out.stdout        # tests/cpp/test_cpu_ops.cc: CPU operators + serializer on libtorch
