"""edward.inferences, restricted to the HMC hot path and its stochastic-gradient siblings."""
from .hmc import HMC
from .inference import Inference
from .monte_carlo import MonteCarlo
from .sgmcmc import SGHMC, SGLD

__all__ = ["Inference", "MonteCarlo", "HMC", "SGLD", "SGHMC"]
