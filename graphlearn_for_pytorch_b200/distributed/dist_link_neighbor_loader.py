from .dist_loader import DistLinkNeighborLoader  # noqa: F401
