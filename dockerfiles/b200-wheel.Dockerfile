# Wheel builder (counterpart of the reference's dockerfiles/graphlearn-torch-wheel.Dockerfile).
# CUDA 12.9 toolkit (first nvcc with sm_100a/sm_103a), the torch the wheel is linked against, nothing else.
#   docker build -f dockerfiles/b200-wheel.Dockerfile -t glt-b200-wheel .
#   docker run --rm -v $PWD:/src -w /src glt-b200-wheel scripts/build_wheel.sh dist
FROM nvidia/cuda:12.9.0-devel-ubuntu24.04
RUN apt-get update && apt-get install -y --no-install-recommends python3 python3-pip python3-venv python3-dev git ninja-build \
    && rm -rf /var/lib/apt/lists/*
RUN python3 -m venv /opt/venv
ENV PATH=/opt/venv/bin:$PATH CUDA_HOME=/usr/local/cuda TORCH_CUDA_ARCH_LIST=10.0a
ARG TORCH_VERSION=2.11.0
RUN pip install --no-cache-dir "torch==${TORCH_VERSION}" --index-url https://download.pytorch.org/whl/cu128 \
    && pip install --no-cache-dir numpy setuptools wheel ninja pybind11 pytest
WORKDIR /src
CMD ["scripts/build_wheel.sh", "dist"]
