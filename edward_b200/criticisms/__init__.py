"""edward.criticisms for the models of this path: evaluate (on the device sample stores) and ppc."""
from .evaluate import evaluate
from .ppc import ppc

__all__ = ["evaluate", "ppc"]
