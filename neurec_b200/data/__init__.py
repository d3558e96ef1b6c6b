"""Host-side mirror of the reference's `data` package (data/__init__.py:1-4)."""
from .dataset import Dataset
from .sampler import PairwiseSampler, PointwiseSampler
from .sampler import TimeOrderPairwiseSampler, TimeOrderPointwiseSampler
