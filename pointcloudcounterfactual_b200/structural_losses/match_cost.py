"""``match_cost`` operator (approximate-matching EMD), mirroring the reference's
external/pytorch_structural_losses/structural_losses/match_cost.py:11-50.

The reference's forward runs ApproxMatch + MatchCost and pins the (B,M,N) match matrix (512 MiB at B=32, N=2048)
on ``ctx`` until backward runs MatchCostGrad.  Here the forward is ONE fused call that returns the cost and the
unit-gradient tensors; backward only scales them by grad_output (match_cost.py:41-42) -- O(B(N+M)) memory.
"""
from __future__ import annotations

from typing import Any

import torch
from torch.autograd import Function

from .structural_losses_backend import MatchCostFused


class MatchCostFunction(Function):
    """(set1 (B,N,3), set2 (B,M,3)) -> cost (B,)."""

    @staticmethod
    def forward(ctx: Any, set1: torch.Tensor, set2: torch.Tensor) -> torch.Tensor:
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        cost, grad1, grad2 = MatchCostFused(set1, set2, want_grad1=need1, want_grad2=need2)
        ctx.unit_grads = (grad1, grad2)
        return cost

    @staticmethod
    def backward(ctx: Any, grad_output: torch.Tensor):
        grad1, grad2 = ctx.unit_grads
        scale = grad_output.reshape(-1, 1, 1)
        return (None if grad1 is None else grad1 * scale), (None if grad2 is None else grad2 * scale)


def match_cost(set1: torch.Tensor, set2: torch.Tensor) -> torch.Tensor:
    return MatchCostFunction.apply(set1, set2)
