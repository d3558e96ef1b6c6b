"""Host-side mirror of the reference's `model` package for the MF family."""
from .AbstractRecommender import AbstractRecommender
