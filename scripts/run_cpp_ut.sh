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
if [[ "${1:-}" == "tsan" ]]; then
  g++ -std=c++17 -O1 -g -fsanitize=thread -I graphlearn_for_pytorch_b200/csrc $SRC -o "$OUT/test_shm_queue_tsan" -lpthread -lrt
  TSAN_OPTIONS="halt_on_error=1" "$OUT/test_shm_queue_tsan" --threads-only
fi
