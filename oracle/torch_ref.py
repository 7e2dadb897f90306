"""Restatement of the reference's torch CPU path -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

This is "the reference torch CPU path" BASELINE.json config 1 names: what the reference runs when no GPU is
present (``src/utils/neighbour_ops.py:29,65`` dispatch on ``device.type``).  Same operations in the same order
(dense GEMM-form distances, ``min`` / ``topk``), so its timing is the reference's timing.
Pinned by tests/test_oracle_golden.py::test_torch_ref_matches_reference against tests/golden/*.npz, which were
produced by the genuine reference functions.
"""
from __future__ import annotations

import torch


def torch_square_distance(t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor:
    """src/utils/neighbour_ops.py:43-50: -2 t1 t2^T + |t1|^2 + |t2|^2, (B,N,D) x (B,M,D) -> (B,N,M)."""
    t2t = t2.transpose(-1, -2)
    dist = -2 * torch.matmul(t1, t2t)
    dist += torch.sum(t1 ** 2, -1, keepdim=True)
    dist += torch.sum(t2t ** 2, -2, keepdim=True)
    return dist


def torch_chamfer(t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor:
    """src/train/metrics_and_losses.py:44-47: sum of the row minima plus sum of the column minima."""
    dist = torch_square_distance(t1, t2)
    return torch.min(dist, dim=-1)[0].sum(1) + torch.min(dist, dim=-2)[0].sum(1)


def self_square_distance(t1: torch.Tensor) -> torch.Tensor:
    """src/utils/neighbour_ops.py:53-60 for channels-first (B,C,N)."""
    t2 = t1.transpose(-1, -2)
    sq = torch.sum(t1 ** 2, -2, keepdim=True)
    dist = torch.tensor(-2) * torch.matmul(t2, t1)
    dist += sq
    dist += sq.transpose(-1, -2)
    return dist


def torch_knn(x: torch.Tensor, k: int) -> torch.Tensor:
    """src/utils/neighbour_ops.py:71-74."""
    return self_square_distance(x).topk(k=k, largest=False)[1]


def chamfer_fwd_bwd(t1: torch.Tensor, t2: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """One training-style evaluation: loss per cloud and d(sum loss)/d t1 through autograd."""
    t1 = t1.detach().requires_grad_(True)
    loss = torch_chamfer(t1, t2)
    loss.sum().backward()
    return loss.detach(), t1.grad
