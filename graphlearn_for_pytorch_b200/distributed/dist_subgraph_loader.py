from .dist_loader import DistSubGraphLoader  # noqa: F401
