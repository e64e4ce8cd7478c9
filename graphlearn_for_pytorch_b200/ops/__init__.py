"""Loader of the native core.

``native`` is the pybind module built from ``csrc/`` (see ``ops/build.py``).  It is
loaded straight from the in-tree shared object -- never through a JIT cache -- so
the GPU box runs exactly the binary that was built with the sources.  When a GPU
is present and the extension is missing we fail loudly instead of silently
falling back to eager PyTorch.
"""
import importlib.util
import os
import sys

import torch  # noqa: F401  (must be imported before the extension: libtorch symbols)

from . import build as _build

native = None
_load_error = None


def _load():
  global native, _load_error
  if native is not None:
    return native
  path = _build.SO_PATH
  try:
    if _build.is_stale() and os.environ.get("GLT_B200_NO_AUTOBUILD", "0") != "1":
      if not os.path.exists(path) or os.environ.get("GLT_B200_REBUILD", "0") == "1":
        _build.build()
    spec = importlib.util.spec_from_file_location(_build.EXT_NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[_build.EXT_NAME] = mod
    native = mod
  except Exception as e:  # pragma: no cover - build/toolchain problems
    _load_error = e
    native = None
  return native


def require_native():
  """Return the native module or raise with the original load error."""
  mod = _load()
  if mod is None:
    raise RuntimeError(
      "graphlearn_for_pytorch_b200 native extension is not available "
      f"({_load_error!r}); run `python -m graphlearn_for_pytorch_b200.ops.build`")
  return mod


def has_native() -> bool:
  return _load() is not None


def cuda_available() -> bool:
  return torch.cuda.is_available() and has_native()


_load()
