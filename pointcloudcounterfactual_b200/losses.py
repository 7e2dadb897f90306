"""Chamfer reconstruction losses with the reference's names and reductions
(``src/train/metrics_and_losses.py:21-47``), computed by the sm_100a ``nn_distance`` operator.

The reference's two back-ends disagree on the reduction: ``pykeops_chamfer`` (the one
used in GPU training) takes the MEAN over points, ``torch_chamfer`` the SUM.  Both are kept as they are.
"""
from __future__ import annotations

import torch

from .structural_losses import match_cost, nn_distance


def pykeops_chamfer(t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor:
    """(:21-41) (B,N,3),(B,M,3) -> (B,): mean_j min_k |t1_j - t2_k|^2 + mean_k min_j |t2_k - t1_j|^2.
    Gradients reach both clouds through the nearest-neighbour pairs, like the reference's gather-based form."""
    dist1, dist2 = nn_distance(t1.contiguous(), t2.contiguous())
    return dist2.mean(1) + dist1.mean(1)


def torch_chamfer(t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor:
    """(:44-47) same pairs, SUM over points."""
    dist1, dist2 = nn_distance(t1.contiguous(), t2.contiguous())
    return dist1.sum(1) + dist2.sum(1)


def chamfer_emd(recon: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """The reference's ChamferEMD reconstruction loss on GPU (:70-79 -> :59-67 + :50-56):
    ``pykeops_chamfer(recon, ref) + match_cost(recon, ref)``, one value per cloud."""
    return pykeops_chamfer(recon, ref) + match_cost(recon.contiguous(), ref.contiguous())
