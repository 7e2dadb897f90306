"""Drop-in for the reference package ``structural_losses`` (external/pytorch_structural_losses/structural_losses/
__init__.py:1-5): exports ``match_cost`` and ``nn_distance`` backed by the sm_100a kernels."""
from .match_cost import match_cost
from .nn_distance import nn_distance

__all__ = ["match_cost", "nn_distance"]
