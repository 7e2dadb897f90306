"""Drop-in for the reference's pybind module ``structural_losses.structural_losses_backend``
(external/pytorch_structural_losses/src/structural_loss.cpp:129-135): same five function names, argument order,
output shapes/dtypes and ownership (outputs are allocated here, uninitialised, on the inputs' device and fully
written by the kernels).  Each function is a thin call into the C ABI of include/pcc_b200.h on the current stream.
"""
from __future__ import annotations

import torch

from .. import _lib as L


def _dims(set_d: torch.Tensor, set_q: torch.Tensor) -> tuple[int, int, int]:
    if set_d.dim() != 3 or set_q.dim() != 3 or set_d.size(2) != 3 or set_q.size(2) != 3:
        raise RuntimeError("expected point sets of shape (batch, points, 3)")
    if set_d.size(0) != set_q.size(0):
        raise RuntimeError("batch sizes differ")
    return set_d.size(0), set_d.size(1), set_q.size(1)


def NNDistance(set_d: torch.Tensor, set_q: torch.Tensor) -> list[torch.Tensor]:
    """-> [dist1 (B,N) f32, idx1 (B,N) i32, dist2 (B,M) f32, idx2 (B,M) i32]  (structural_loss.cpp:81-100)."""
    L.require_cuda(set_d, set_q)
    b, n, m = _dims(set_d, set_q)
    with torch.cuda.device(set_d.device):
        dist1 = torch.empty((b, n), dtype=torch.float32, device=set_d.device)
        idx1 = torch.empty((b, n), dtype=torch.int32, device=set_d.device)
        dist2 = torch.empty((b, m), dtype=torch.float32, device=set_d.device)
        idx2 = torch.empty((b, m), dtype=torch.int32, device=set_d.device)
        L.check(L.load().pcc_nndistance(b, n, L.ptr(set_d), m, L.ptr(set_q), L.ptr(dist1), L.ptr(idx1), L.ptr(dist2),
                                        L.ptr(idx2), L.stream_of(set_d)), "NNDistance")
    return [dist1, idx1, dist2, idx2]


def NNDistanceTC(set_d: torch.Tensor, set_q: torch.Tensor) -> list[torch.Tensor]:
    """B200 addition: NNDistance through the tcgen05 tensor cores (``pcc_nndistance_tc``, csrc/chamfer_tc.cu); same outputs,
    bit-identical values; raises for clouds outside 256..2560 points."""
    L.require_cuda(set_d, set_q)
    b, n, m = _dims(set_d, set_q)
    with torch.cuda.device(set_d.device):
        dist1 = torch.empty((b, n), dtype=torch.float32, device=set_d.device)
        idx1 = torch.empty((b, n), dtype=torch.int32, device=set_d.device)
        dist2 = torch.empty((b, m), dtype=torch.float32, device=set_d.device)
        idx2 = torch.empty((b, m), dtype=torch.int32, device=set_d.device)
        L.check(L.load().pcc_nndistance_tc(b, n, L.ptr(set_d), m, L.ptr(set_q), L.ptr(dist1), L.ptr(idx1), L.ptr(dist2),
                                           L.ptr(idx2), L.stream_of(set_d)), "NNDistanceTC")
    return [dist1, idx1, dist2, idx2]


def NNDistanceGrad(set_d: torch.Tensor, set_q: torch.Tensor, idx1: torch.Tensor, idx2: torch.Tensor,
                   grad_dist1: torch.Tensor, grad_dist2: torch.Tensor) -> list[torch.Tensor]:
    """-> [grad1 (B,N,3), grad2 (B,M,3)]  (structural_loss.cpp:102-125)."""
    L.require_cuda(set_d, set_q, grad_dist1, grad_dist2)
    L.require_cuda(idx1, idx2, dtype=torch.int32)
    b, n, m = _dims(set_d, set_q)
    with torch.cuda.device(set_d.device):
        grad1 = torch.empty((b, n, 3), dtype=torch.float32, device=set_d.device)
        grad2 = torch.empty((b, m, 3), dtype=torch.float32, device=set_d.device)
        L.check(L.load().pcc_nndistancegrad(b, n, L.ptr(set_d), m, L.ptr(set_q), L.ptr(grad_dist1), L.ptr(idx1),
                                            L.ptr(grad_dist2), L.ptr(idx2), L.ptr(grad1), L.ptr(grad2),
                                            L.stream_of(set_d)), "NNDistanceGrad")
    return [grad1, grad2]


def ApproxMatch(set_d: torch.Tensor, set_q: torch.Tensor) -> list[torch.Tensor]:
    """-> [match (B,M,N), temp (B,2(N+M))]  (structural_loss.cpp:24-38)."""
    L.require_cuda(set_d, set_q)
    b, n, m = _dims(set_d, set_q)
    with torch.cuda.device(set_d.device):
        match = torch.empty((b, m, n), dtype=torch.float32, device=set_d.device)
        temp = torch.empty((b, (n + m) * 2), dtype=torch.float32, device=set_d.device)
        L.check(L.load().pcc_approxmatch(b, n, m, L.ptr(set_d), L.ptr(set_q), L.ptr(match), L.ptr(temp),
                                         L.stream_of(set_d)), "ApproxMatch")
    return [match, temp]


def MatchCost(set_d: torch.Tensor, set_q: torch.Tensor, match: torch.Tensor) -> torch.Tensor:
    """-> cost (B,)  (structural_loss.cpp:40-53)."""
    L.require_cuda(set_d, set_q, match)
    b, n, m = _dims(set_d, set_q)
    if tuple(match.shape) != (b, m, n):
        raise RuntimeError("match must have shape (batch, #query_points, #dataset_points)")
    with torch.cuda.device(set_d.device):
        out = torch.empty((b,), dtype=torch.float32, device=set_d.device)
        L.check(L.load().pcc_matchcost(b, n, m, L.ptr(set_d), L.ptr(set_q), L.ptr(match), L.ptr(out),
                                       L.stream_of(set_d)), "MatchCost")
    return out


def MatchCostGrad(set_d: torch.Tensor, set_q: torch.Tensor, match: torch.Tensor) -> list[torch.Tensor]:
    """-> [grad1 (B,N,3), grad2 (B,M,3)]  (structural_loss.cpp:55-70)."""
    L.require_cuda(set_d, set_q, match)
    b, n, m = _dims(set_d, set_q)
    if tuple(match.shape) != (b, m, n):
        raise RuntimeError("match must have shape (batch, #query_points, #dataset_points)")
    with torch.cuda.device(set_d.device):
        grad1 = torch.empty((b, n, 3), dtype=torch.float32, device=set_d.device)
        grad2 = torch.empty((b, m, 3), dtype=torch.float32, device=set_d.device)
        L.check(L.load().pcc_matchcostgrad(b, n, m, L.ptr(set_d), L.ptr(set_q), L.ptr(match), L.ptr(grad1),
                                           L.ptr(grad2), L.stream_of(set_d)), "MatchCostGrad")
    return [grad1, grad2]


def MatchCostFused(set_d: torch.Tensor, set_q: torch.Tensor, want_grad1: bool = True,
                   want_grad2: bool = True) -> tuple[torch.Tensor, torch.Tensor | None, torch.Tensor | None]:
    """B200 addition: ApproxMatch -> MatchCost -> MatchCostGrad in one call without the (B,M,N) matrix.

    -> (cost (B,), grad1 (B,N,3) | None, grad2 (B,M,3) | None); gradients are for unit upstream gradient.
    """
    L.require_cuda(set_d, set_q)
    b, n, m = _dims(set_d, set_q)
    with torch.cuda.device(set_d.device):
        cost = torch.empty((b,), dtype=torch.float32, device=set_d.device)
        grad1 = torch.empty((b, n, 3), dtype=torch.float32, device=set_d.device) if want_grad1 else None
        grad2 = torch.empty((b, m, 3), dtype=torch.float32, device=set_d.device) if want_grad2 else None
        temp = torch.empty((b, (n + m) * 2), dtype=torch.float32, device=set_d.device)
        L.check(L.load().pcc_matchcost_fused(b, n, m, L.ptr(set_d), L.ptr(set_q), L.ptr(cost), L.ptr(grad1),
                                             L.ptr(grad2), L.ptr(temp), L.stream_of(set_d)), "MatchCostFused")
    return cost, grad1, grad2
