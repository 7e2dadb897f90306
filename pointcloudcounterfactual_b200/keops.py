"""A KeOps-free stand-in for the four ``pykeops.torch.LazyTensor`` patterns the reference uses:

    d = ((LazyTensor(t1[:, :, None, :]) - LazyTensor(t2[:, None, :, :])) ** 2).sum(-1)   neighbour_ops.py:37-40
    d.argKmin(k, dim=2)  -> (B, N, k) int64                                              neighbour_ops.py:81
    d.argmin(axis=1) / d.argmin(axis=2) -> (B, M, 1) / (B, N, 1) int64       metrics_and_losses.py:33,36; quantize.py:28
    d.sum(1) -> (B, M, 1)                                                                quantize.py:31

plus ``pykeops.set_verbose`` (neighbour_ops.py:13).  The arg-reductions run the sm_100a argKmin kernel
(``pcc_argkmin``); nothing is materialised.  ``install()`` registers this module as ``pykeops`` /
``pykeops.torch`` so the reference's ``src/`` runs unchanged without PyKeOps.
"""
from __future__ import annotations

import sys
import types

import torch

from . import _lib as L


def _same_view(a: torch.Tensor, b: torch.Tensor) -> bool:
    return (a.data_ptr() == b.data_ptr() and a.shape == b.shape and a.stride() == b.stride() and a.dtype == b.dtype
            and a.device == b.device)


def argkmin(q: torch.Tensor, r: torch.Tensor, k: int, return_dist: bool = False):
    """q (B,Nq,C), r (B,Nr,C) CUDA fp32 -> idx (B,Nq,k) int64, ascending by (squared distance, index).

    When both operands are the same storage -- the reference's ``pykeops_knn`` builds its expression on (x, x),
    neighbour_ops.py:79-81 -- the C entry point sees q == r and runs the point-major SELF kNN route: the
    warp-cooperative xyz kernel for C == 3, the tcgen05 kernels for C % 32 == 0 (``pcc_argkmin``, csrc/knn.cu)."""
    same = _same_view(q, r)
    q = q.contiguous()
    r = q if same else r.contiguous()
    L.require_cuda(q, r)
    if q.dim() != 3 or r.dim() != 3 or q.size(0) != r.size(0) or q.size(2) != r.size(2):
        raise RuntimeError("argkmin expects (B,Nq,C) and (B,Nr,C)")
    b, nq, c = q.shape
    nr = r.size(1)
    with torch.cuda.device(q.device):
        idx = torch.empty((b, nq, k), dtype=torch.int64, device=q.device)
        dist = torch.empty((b, nq, k), dtype=torch.float32, device=q.device) if return_dist else None
        L.check(L.load().pcc_argkmin(b, nq, nr, c, k, L.ptr(q), L.ptr(r), L.ptr(idx), L.ptr(dist), L.stream_of(q)),
                "argkmin")
    return (idx, dist) if return_dist else idx


class LazyTensor:
    """Symbolic (B,N,1,D) or (B,1,M,D) operand."""

    def __init__(self, t: torch.Tensor):
        if t.dim() != 4 or (t.size(1) != 1 and t.size(2) != 1):
            raise NotImplementedError("LazyTensor shim: expected t[:, :, None, :] or t[:, None, :, :]")
        self.t = t
        self.axis = 1 if t.size(2) == 1 else 2  # which symbolic axis this operand indexes (i -> 1, j -> 2)
        if t.size(1) == 1 and t.size(2) == 1:
            self.axis = 0  # a single point broadcasts on both axes; resolved by the partner operand

    def __sub__(self, other: "LazyTensor") -> "_Diff":
        return _Diff(self, other)


class _Diff:
    def __init__(self, a: LazyTensor, b: LazyTensor):
        ax_a, ax_b = a.axis, b.axis
        if ax_a == 0:
            ax_a = 3 - ax_b if ax_b else 1
        if ax_b == 0:
            ax_b = 3 - ax_a
        if {ax_a, ax_b} != {1, 2}:
            raise NotImplementedError("LazyTensor shim: operands must index different axes")
        i_op, j_op = (a, b) if ax_a == 1 else (b, a)
        self.ti = i_op.t.squeeze(2)  # (B,N,D)
        self.tj = j_op.t.squeeze(1)  # (B,M,D)

    def __pow__(self, p: int) -> "_SqDiff":
        if p != 2:
            raise NotImplementedError("LazyTensor shim: only **2")
        return _SqDiff(self.ti, self.tj)


class _SqDiff:
    def __init__(self, ti: torch.Tensor, tj: torch.Tensor):
        self.ti, self.tj = ti, tj

    def sum(self, dim: int = -1) -> "SquareDistance":
        if dim not in (-1, 3):
            raise NotImplementedError("LazyTensor shim: the feature axis must be reduced first")
        return SquareDistance(self.ti, self.tj)


class SquareDistance:
    """Symbolic (B,N,M) matrix of squared distances between ti (B,N,D) and tj (B,M,D)."""

    def __init__(self, ti: torch.Tensor, tj: torch.Tensor):
        self.ti, self.tj = ti, tj
        self._nn = None  # (versions, dist1, idx1, dist2, idx2) of one fused Chamfer launch

    def _nn_pair(self):
        """xyz clouds, k = 1 (``pykeops_chamfer``, metrics_and_losses.py:32-36: argmin over axis 1, then over axis 2, of ONE
        expression): both directions come from ONE ``pcc_nndistance`` launch, kept for the second reduction.  Distances
        and the lowest-index tie rule are those of the reference's own Chamfer kernel (nndistance.cu:2-124)."""
        ti, tj = self.ti, self.tj
        ok = (ti.is_cuda and tj.is_cuda and ti.dtype == torch.float32 and tj.dtype == torch.float32 and ti.dim() == 3
              and tj.dim() == 3 and ti.size(2) == 3 and tj.size(2) == 3 and ti.size(0) == tj.size(0) and ti.size(0) > 0
              and ti.size(1) > 64 and tj.size(1) > 64)  # tiny problems (vector quantisation) keep the argKmin route
        if not ok:
            return None
        ver = (ti._version, tj._version)
        if self._nn is None or self._nn[0] != ver:
            from .structural_losses.structural_losses_backend import NNDistance

            d1, i1, d2, i2 = NNDistance(ti.detach().contiguous(), tj.detach().contiguous())
            self._nn = (ver, d1, i1, d2, i2)
        return self._nn[1:]

    def argKmin(self, K: int, dim: int = 2, axis: int | None = None) -> torch.Tensor:
        dim = dim if axis is None else axis
        if dim == 2:
            return argkmin(self.ti, self.tj, K)
        if dim == 1:
            return argkmin(self.tj, self.ti, K)
        raise NotImplementedError("argKmin over dim 1 or 2 only")

    def argmin(self, axis: int | None = None, dim: int | None = None) -> torch.Tensor:
        d = axis if axis is not None else dim
        nn = self._nn_pair() if d in (1, 2) else None
        if nn is not None:  # axis 2: nearest tj for every ti (idx1); axis 1: nearest ti for every tj (idx2)
            return (nn[1] if d == 2 else nn[3]).to(torch.int64).unsqueeze(-1)
        return self.argKmin(1, dim=d)

    def _dense(self) -> torch.Tensor:
        # only quantize.py's tiny (B*n_codes, 1, 16) case reaches this; dense torch keeps it differentiable (the
        # arg-reductions of the same expression run argmin_small_kernel / the argKmin kernels)
        diff = self.ti[:, :, None, :] - self.tj[:, None, :, :]
        return (diff * diff).sum(-1)

    def sum(self, dim: int) -> torch.Tensor:
        if dim == 1:
            return self._dense().sum(1).unsqueeze(-1)
        if dim == 2:
            return self._dense().sum(2).unsqueeze(-1)
        raise NotImplementedError("sum over dim 1 or 2 only")

    def min(self, axis: int | None = None, dim: int | None = None) -> torch.Tensor:
        d = axis if axis is not None else dim
        nn = self._nn_pair() if d in (1, 2) else None
        if nn is not None:
            return (nn[0] if d == 2 else nn[2]).unsqueeze(-1)
        q, r = (self.ti, self.tj) if d == 2 else (self.tj, self.ti)
        return argkmin(q, r, 1, return_dist=True)[1]


def set_verbose(_flag: bool = False) -> None:
    """pykeops.set_verbose stand-in (neighbour_ops.py:13): nothing is JIT-compiled here."""


def install() -> None:
    """Register this shim as ``pykeops`` and ``pykeops.torch`` (only if the real PyKeOps is not already imported)."""
    if "pykeops" in sys.modules and not getattr(sys.modules["pykeops"], "_pcc_b200_shim", False):
        return
    pk = types.ModuleType("pykeops")
    pk._pcc_b200_shim = True
    pk.set_verbose = set_verbose
    pkt = types.ModuleType("pykeops.torch")
    pkt.LazyTensor = LazyTensor
    pk.torch = pkt
    sys.modules["pykeops"] = pk
    sys.modules["pykeops.torch"] = pkt
