"""edward_b200 — Edward's HMC hot path (ed.HMC over ed.models.Empirical for GLM-style models) rebuilt
for B200: hand-written sm_100a CUDA kernels behind a C ABI (libedhmc.so), driven through the reference's
own Python surface.

    import edward_b200 as ed
    from edward_b200 import tfshim as tf
    from edward_b200.models import Bernoulli, Empirical, Normal
"""
from . import criticisms, inferences, models, util  # noqa: F401
from .criticisms import evaluate, ppc  # noqa: F401
from .util.copying import copy  # noqa: F401
from .inferences import HMC, SGHMC, SGLD, Inference, MonteCarlo  # noqa: F401
from .models import RandomVariable  # noqa: F401
from .util import Progbar, check_data, check_latent_vars, dot, get_session, random_variables, set_seed, transform  # noqa: F401

__version__ = "0.1.0"
