"""edward.models, restricted to the HMC hot path (edward/models/__init__.py:17-29)."""
from .empirical import Empirical
from .random_variable import RandomVariable
from .random_variables import Bernoulli, Beta, Normal, Poisson, TransformedDistribution

__all__ = ["RandomVariable", "Empirical", "Normal", "Bernoulli", "Poisson", "Beta", "TransformedDistribution"]
