from .base import *
from .negative_sampler import RandomNegativeSampler
from .neighbor_sampler import NeighborSampler
