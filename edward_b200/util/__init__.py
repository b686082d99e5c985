"""edward.util, restricted to the HMC hot path (edward/util/__init__.py:15-38)."""
from .graphs import get_seed, get_session, random_variables, set_seed
from .progbar import Progbar
from .random_variables import check_data, check_latent_vars, transform
from .tensorflow import dot

__all__ = ["check_data", "check_latent_vars", "dot", "get_session", "Progbar", "random_variables", "set_seed", "transform"]
