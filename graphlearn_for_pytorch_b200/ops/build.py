"""In-tree build of the native core (C++ CPU ops + sm_100a CUDA kernels).

The reference builds through CMake / setup.py for sm_50..sm_75 generic SIMT code
(reference: graphlearn_torch/python/utils/build_glt.py:71-135).  Here every CUDA
source is compiled for exactly one target, ``sm_100a``, and the resulting shared
object lives inside the package (``_ext/glt_b200_C.so``) so it travels with the
source tree.
"""
import glob
import os
import sys

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG_DIR, "csrc")
EXT_DIR = os.path.join(PKG_DIR, "_ext")
EXT_NAME = "glt_b200_C"
SO_PATH = os.path.join(EXT_DIR, EXT_NAME + ".so")

NVCC_FLAGS = [
  "-gencode", "arch=compute_100a,code=sm_100a",
  "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
]
CXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-Wno-sign-compare"]


def sources():
  src = [os.path.join(CSRC, "bindings.cc")]
  src += sorted(glob.glob(os.path.join(CSRC, "cpu", "*.cc")))
  src += sorted(glob.glob(os.path.join(CSRC, "cuda", "*.cu")))
  return src


def is_stale() -> bool:
  if not os.path.exists(SO_PATH):
    return True
  so_mtime = os.path.getmtime(SO_PATH)
  deps = sources() + glob.glob(os.path.join(CSRC, "**", "*.h"), recursive=True) \
      + glob.glob(os.path.join(CSRC, "**", "*.cuh"), recursive=True)
  return any(os.path.getmtime(p) > so_mtime for p in deps)


def build(verbose: bool = False, force: bool = False) -> str:
  """Compile (if needed) and return the path of the shared object."""
  if not force and not is_stale():
    return SO_PATH
  from torch.utils import cpp_extension
  os.makedirs(EXT_DIR, exist_ok=True)
  os.environ.setdefault("MAX_JOBS", str(min(8, os.cpu_count() or 4)))
  cpp_extension.load(
    name=EXT_NAME,
    sources=sources(),
    extra_cflags=CXX_FLAGS,
    extra_cuda_cflags=NVCC_FLAGS,
    extra_ldflags=["-lrt", "-lpthread"],
    build_directory=EXT_DIR,
    with_cuda=True,
    is_python_module=False,
    verbose=verbose,
  )
  if not os.path.exists(SO_PATH):
    raise RuntimeError(f"native build did not produce {SO_PATH}")
  return SO_PATH


if __name__ == "__main__":
  print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
