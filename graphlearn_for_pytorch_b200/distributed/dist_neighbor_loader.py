from .dist_loader import DistNeighborLoader  # noqa: F401
