# Development image for graphlearn_for_pytorch_b200 (counterpart of the reference's dockerfiles/graphlearn-torch-dev.Dockerfile).
# CUDA 12.9 is the first toolkit whose nvcc accepts -gencode arch=compute_100a,code=sm_100a.
FROM nvidia/cuda:12.9.0-devel-ubuntu24.04
ENV DEBIAN_FRONTEND=noninteractive
RUN apt-get update && apt-get install -y --no-install-recommends python3 python3-pip python3-venv git ninja-build g++ \
    openssh-client && rm -rf /var/lib/apt/lists/*
RUN python3 -m venv /opt/venv
ENV PATH=/opt/venv/bin:$PATH CUDA_HOME=/usr/local/cuda
RUN pip install --no-cache-dir torch --index-url https://download.pytorch.org/whl/cu128 && \
    pip install --no-cache-dir numpy pyarrow pyyaml paramiko pytest pytest-timeout pybind11 ninja
WORKDIR /workspace/graphlearn_for_pytorch_b200
COPY . .
# nvcc cross-compiles sm_100a without a GPU: the image is built on ordinary CI runners
RUN python -c "import __graft_entry__ as g; g.build()"
CMD ["bash", "scripts/run_tests.sh"]
