"""``nn_distance`` operator (Chamfer nearest-neighbour distances), mirroring the reference's
external/pytorch_structural_losses/structural_losses/nn_distance.py:9-43."""
from __future__ import annotations

from typing import Any

import torch
from torch.autograd import Function

from .structural_losses_backend import NNDistance, NNDistanceGrad


class NNDistanceFunction(Function):
    """(set1 (B,N,3), set2 (B,M,3)) -> (dist1 (B,N), dist2 (B,M)); nearest indices are kept for backward."""

    @staticmethod
    def forward(ctx: Any, set1: torch.Tensor, set2: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        dist1, idx1, dist2, idx2 = NNDistance(set1, set2)
        ctx.save_for_backward(set1, set2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx: Any, grad_dist1: torch.Tensor, grad_dist2: torch.Tensor):
        set1, set2, idx1, idx2 = ctx.saved_tensors
        grad1, grad2 = NNDistanceGrad(set1, set2, idx1, idx2, grad_dist1.contiguous(), grad_dist2.contiguous())
        return grad1, grad2


def nn_distance(set1: torch.Tensor, set2: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    return NNDistanceFunction.apply(set1, set2)
