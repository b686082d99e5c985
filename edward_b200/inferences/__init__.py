"""edward.inferences, restricted to the HMC hot path."""
from .hmc import HMC
from .inference import Inference
from .monte_carlo import MonteCarlo

__all__ = ["Inference", "MonteCarlo", "HMC"]
