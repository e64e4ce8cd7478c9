"""Minimal stand-in for the `torch_geometric.data` containers imported by the UNMODIFIED
reference (loader/transform.py:20, distributed/dist_loader.py:19).  PyG is not installable
offline; only attribute-bag behaviour is needed by the reference code paths we run."""
from . import data  # noqa: F401
