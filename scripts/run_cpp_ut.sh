#!/usr/bin/env bash
# Native (C++) unit tests that need neither Python nor a GPU (the reference's scripts/run_cpp_ut.sh runs its
# GoogleTest binaries; gtest is not in this image, the tests use plain CHECKs and exit codes).
#   scripts/run_cpp_ut.sh          normal build
#   scripts/run_cpp_ut.sh tsan     + ThreadSanitizer build of the in-process producer/consumer test
set -euo pipefail
cd "$(dirname "$0")/.."
OUT=${TMPDIR:-/tmp}/glt_b200_cpp_ut
mkdir -p "$OUT"
SRC="tests/cpp/test_shm_queue.cc graphlearn_for_pytorch_b200/csrc/cpu/shm_queue.cc"
g++ -std=c++17 -O2 -g -I graphlearn_for_pytorch_b200/csrc $SRC -o "$OUT/test_shm_queue" -lpthread -lrt
"$OUT/test_shm_queue"
# CPU operators + serializer against libtorch (no Python interpreter involved)
PY=${PYTHON:-python}
TORCH_INC=$($PY -c "from torch.utils.cpp_extension import include_paths; print(' '.join('-isystem ' + p for p in include_paths()))")
TORCH_LIB=$($PY -c "from torch.utils.cpp_extension import library_paths; print(library_paths()[0])")
PY_INC=$($PY -c "import sysconfig; print(sysconfig.get_paths()['include'])")
# the op headers include <torch/extension.h> (pybind11): link libpython, although no interpreter is started
PY_LD=$($PY -c "import sysconfig; print('-L' + sysconfig.get_config_var('LIBDIR') + ' -lpython' + sysconfig.get_config_var('LDVERSION'))")
EXT=$(pwd)/graphlearn_for_pytorch_b200/_ext
if [[ -f "$EXT/glt_b200_C.so" ]]; then
  # link the test against the SHIPPED extension: what is tested is the binary that Python loads
  g++ -std=c++17 -O0 -g -fPIC -I graphlearn_for_pytorch_b200/csrc $TORCH_INC -isystem "$PY_INC" \
      -isystem "${CUDA_HOME:-/usr/local/cuda}/include" tests/cpp/test_cpu_ops.cc -o "$OUT/test_cpu_ops" \
      -L"$EXT" -l:glt_b200_C.so -Wl,-rpath,"$EXT" -L"$TORCH_LIB" -Wl,-rpath,"$TORCH_LIB" \
      -ltorch_python -ltorch -ltorch_cpu -lc10 $PY_LD -lpthread -lrt
else
  g++ -std=c++17 -O0 -g -fPIC -I graphlearn_for_pytorch_b200/csrc $TORCH_INC -isystem "$PY_INC" \
      -isystem "${CUDA_HOME:-/usr/local/cuda}/include" \
      tests/cpp/test_cpu_ops.cc graphlearn_for_pytorch_b200/csrc/cpu/cpu_ops.cc \
      graphlearn_for_pytorch_b200/csrc/cpu/sample_queue.cc graphlearn_for_pytorch_b200/csrc/cpu/shm_queue.cc \
      -o "$OUT/test_cpu_ops" -L"$TORCH_LIB" -Wl,-rpath,"$TORCH_LIB" -ltorch -ltorch_cpu -lc10 \
      -L"${CUDA_HOME:-/usr/local/cuda}/lib64" -lcudart $PY_LD -lpthread -lrt
fi
LD_LIBRARY_PATH="$TORCH_LIB:${CUDA_HOME:-/usr/local/cuda}/lib64:${LD_LIBRARY_PATH:-}" "$OUT/test_cpu_ops"
if [[ "${1:-}" == "tsan" ]]; then
  g++ -std=c++17 -O1 -g -fsanitize=thread -I graphlearn_for_pytorch_b200/csrc $SRC -o "$OUT/test_shm_queue_tsan" -lpthread -lrt
  TSAN_OPTIONS="halt_on_error=1" "$OUT/test_shm_queue_tsan" --threads-only
fi
