"""Has the interpreter started shutting down?  `__del__` / `shutdown()` guards use this to skip RPC and queue
traffic once module globals may already be torn down (capability parity: reference utils/exit_status.py).

Read it through `is_python_exiting()`: `from .exit_status import python_exit_status` would freeze the value that the
flag had at import time.
"""
import atexit
import sys

python_exit_status = False          # kept for API compatibility (module attribute, updated at exit)


class _ExitFlag(object):
  raised = False


def is_python_exiting() -> bool:
  return _ExitFlag.raised or sys.is_finalizing()


@atexit.register
def _mark_exit():
  _ExitFlag.raised = True
  globals()['python_exit_status'] = True
