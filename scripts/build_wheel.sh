#!/bin/bash
# Build a binary wheel that ships the sm_100a extension (counterpart of the reference's scripts/build_wheel.sh,
# which loops over python versions inside a manylinux container; here the wheel is tied to the interpreter and the
# torch ABI it was built against, so one wheel per (python, torch) pair: run this inside dockerfiles/b200-wheel.Dockerfile).
#   scripts/build_wheel.sh [outdir]      -> outdir/graphlearn_for_pytorch_b200-<ver>-cp3xx-cp3xx-linux_x86_64.whl
set -euo pipefail
cd "$(dirname "$0")/.."
OUT=${1:-dist}
python -m graphlearn_for_pytorch_b200.ops.build -f          # nvcc -gencode arch=compute_100a,code=sm_100a
python - <<'PY'
import graphlearn_for_pytorch_b200 as g
assert g.ops.has_native(), g.ops._load_error
print('native core ok')
PY
python setup.py -q bdist_wheel --dist-dir "$OUT" --plat-name linux_x86_64 --python-tag "cp$(python -c 'import sys; print(f"{sys.version_info[0]}{sys.version_info[1]}")')"
ls -la "$OUT"/*.whl
python - "$OUT" <<'PY'
import glob, sys, zipfile
w = sorted(glob.glob(sys.argv[1] + '/*.whl'))[-1]
names = zipfile.ZipFile(w).namelist()
assert any(n.endswith('_ext/glt_b200_C.so') for n in names), 'extension missing from the wheel'
print('wheel contains', sum(n.endswith('.py') for n in names), 'python files + the sm_100a extension')
PY
