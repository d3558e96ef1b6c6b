"""Host-side mirror of the reference's `data` package (data/__init__.py:1-4).  The TimeOrder*
samplers belong to the sequential models and are outside the hot path."""
from .dataset import Dataset
from .sampler import PairwiseSampler, PointwiseSampler
