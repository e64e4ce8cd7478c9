"""Interpreter-exit flag used by __del__ guards (parity: reference utils/exit_status.py)."""
import atexit

python_exit_status = False


def _set_python_exit_flag():
  global python_exit_status
  python_exit_status = True


atexit.register(_set_python_exit_flag)
