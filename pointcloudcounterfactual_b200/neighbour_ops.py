"""kNN graph construction and the graph operators built on it -- same function names and semantics as the
reference's ``src/utils/neighbour_ops.py`` (line numbers below refer to that file).

``knn`` / ``pykeops_knn`` run the sm_100a kernel (``pcc_knn``): x (B,C,N) channels-first -> (B,N,k) int64, the k
smallest squared distances per point including the point itself, ascending by (distance, index).  There is no CPU
fallback: a CPU tensor raises.  ``torch_knn`` / ``*_square_distance`` keep the reference's dense torch formulas
for callers that ask for them by name.
"""
from __future__ import annotations

import warnings
from typing import Any

import numpy as np
import torch
from torch.autograd import Function

from . import _lib as L
from .keops import LazyTensor, SquareDistance


def knn_indices(x: torch.Tensor, k: int, return_dist: bool = False):
    """x (B,C,N) CUDA fp32 -> idx (B,N,k) int64 [, dist (B,N,k) fp32]."""
    x = x.contiguous()
    L.require_cuda(x)
    if x.dim() != 3:
        raise RuntimeError("knn expects a (batch, channels, points) tensor")
    b, c, n = x.shape
    with torch.cuda.device(x.device):
        idx = torch.empty((b, n, k), dtype=torch.int64, device=x.device)
        dist = torch.empty((b, n, k), dtype=torch.float32, device=x.device) if return_dist else None
        L.check(L.load().pcc_knn(b, c, n, k, L.ptr(x), L.ptr(idx), L.ptr(dist), L.stream_of(x)), "knn")
    return (idx, dist) if return_dist else idx


def index_k_neighbours(pcs, k: int, chunk: int = 2048) -> np.ndarray:
    """Dataset-side kNN (:16-24; the reference builds a scikit-learn KDTree per cloud on the CPU): sequence of (N,3)
    arrays -> (len, N, k) int64 indices, self first, ascending by (distance, index).  Clouds of equal size are batched
    (`chunk` clouds per launch); the caller stores them as int16 under ``index_{k}`` in the h5 file exactly as
    ``src/data/modelnet.py:150-156`` does (``.astype(np.short)``)."""
    pcs = [np.ascontiguousarray(pc, dtype=np.float32) for pc in pcs]
    out: list[np.ndarray | None] = [None] * len(pcs)
    by_size: dict[tuple[int, ...], list[int]] = {}
    for i, pc in enumerate(pcs):
        by_size.setdefault(pc.shape, []).append(i)
    for ids in by_size.values():
        for s0 in range(0, len(ids), chunk):
            sel = ids[s0:s0 + chunk]
            x = torch.from_numpy(np.stack([pcs[i] for i in sel])).cuda().transpose(1, 2).contiguous()  # (B,3,N)
            idx = knn_indices(x, k).cpu().numpy()
            for row, i in enumerate(sel):
                out[i] = idx[row].reshape(-1, k)
    return np.stack(out) if out else np.zeros((0, 0, k), dtype=np.int64)


def square_distance(t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor | SquareDistance:
    """(:27-32) symbolic on CUDA, dense on CPU tensors."""
    if t1.device.type == 'cuda':
        return pykeops_square_distance(t1, t2)
    return torch_square_distance(t1, t2)


def pykeops_square_distance(t1: torch.Tensor, t2: torch.Tensor) -> SquareDistance:
    """(:35-40) symbolic squared-distance matrix between (B,N,D) and (B,M,D)."""
    return ((LazyTensor(t1[:, :, None, :]) - LazyTensor(t2[:, None, :, :])) ** 2).sum(-1)


def torch_square_distance(t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor:
    """(:43-50) dense GEMM-form squared distances, (B,N,D) x (B,M,D) -> (B,N,M)."""
    sq1 = (t1 * t1).sum(-1, keepdim=True)
    sq2 = (t2 * t2).sum(-1).unsqueeze(-2)
    return torch.baddbmm(sq1 + sq2, t1, t2.transpose(-1, -2), alpha=-2.0) if t1.dim() == 3 else (
        sq1 + sq2 - 2.0 * t1 @ t2.transpose(-1, -2))


def self_square_distance(t1: torch.Tensor) -> torch.Tensor:
    """(:53-60) dense self-distances of a channels-first (B,C,N) tensor -> (B,N,N)."""
    sq = (t1 * t1).sum(-2, keepdim=True)
    return sq + sq.transpose(-1, -2) - 2.0 * (t1.transpose(-1, -2) @ t1)


def knn(x: torch.Tensor, k: int) -> torch.Tensor:
    """(:63-68) k nearest neighbours of every point of x (B,C,N), self included."""
    return knn_indices(x, k)


def pykeops_knn(x: torch.Tensor, k: int) -> torch.Tensor:
    """(:77-82)"""
    return knn_indices(x, k)


def torch_knn(x: torch.Tensor, k: int) -> torch.Tensor:
    """(:71-74) dense torch path, kept for callers that name it explicitly."""
    return self_square_distance(x).topk(k=k, largest=False)[1]


_SIDE: dict[int, torch.cuda.Stream] = {}


def _side_stream(device: torch.device) -> torch.cuda.Stream:
    i = device.index if device.index is not None else torch.cuda.current_device()
    if i not in _SIDE:
        _SIDE[i] = torch.cuda.Stream(device)
    return _SIDE[i]


class _GraphGather(Function):
    """x (B,C,N), idx (B,N,k) int64 -> neighbours (B,C,N,k) [mode 0] or graph features (B,2C,N,k) [mode 1] in ONE
    kernel that writes the result once (pcc_graph_gather).  Backward: sums over the edge list sorted by target, no atomics,
    bitwise reproducible (pcc_graph_gather_grad).  When a backward will follow, the sort -- it depends on idx alone -- runs
    during the forward on a side stream, under the HBM-bound gather, and is joined before the forward returns (so the fork
    is safe inside a CUDA graph capture); the backward then consumes it (pcc_graph_gather_grad_presorted, same bits)."""

    @staticmethod
    def forward(ctx: Any, x: torch.Tensor, idx: torch.Tensor, mode: int) -> torch.Tensor:
        b, c, n = x.shape
        k = idx.shape[2]
        lib = L.load()
        ws = None
        with torch.cuda.device(x.device):
            cur = torch.cuda.current_stream(x.device)
            nbytes = int(lib.pcc_graph_edge_sort_bytes(b, n, k)) if ctx.needs_input_grad[0] and idx.is_contiguous() else 0
            if nbytes:
                ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
                side = _side_stream(x.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    L.check(lib.pcc_graph_edge_sort(b, n, k, L.ptr(idx), L.ptr(ws), side.cuda_stream), "graph_edge_sort")
            out = torch.empty((b, (2 * c) if mode else c, n, k), dtype=torch.float32, device=x.device)
            L.check(lib.pcc_graph_gather(b, c, n, k, L.ptr(x), L.ptr(idx), mode, L.ptr(out), L.stream_of(x)), "graph_gather")
            if nbytes:
                cur.wait_stream(side)
        if ws is None:
            ctx.save_for_backward(idx)
        else:
            ctx.save_for_backward(idx, ws)
        ctx.dims = (b, c, n, k, mode)
        return out

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor):
        idx = ctx.saved_tensors[0]
        ws = ctx.saved_tensors[1] if len(ctx.saved_tensors) > 1 else None
        b, c, n, k, mode = ctx.dims
        g = grad_out.contiguous()
        lib = L.load()
        with torch.cuda.device(g.device):
            gx = torch.empty((b, c, n), dtype=torch.float32, device=g.device)
            if ws is not None and g.data_ptr() % 16 == 0:
                L.check(lib.pcc_graph_gather_grad_presorted(b, c, n, k, mode, L.ptr(ws), L.ptr(g), L.ptr(gx), L.stream_of(g)),
                        "graph_gather_grad_presorted")
            else:
                L.check(lib.pcc_graph_gather_grad(b, c, n, k, L.ptr(idx), mode, L.ptr(g), L.ptr(gx), L.stream_of(g)),
                        "graph_gather_grad")
        return gx, None, None


_WARNED: set[str] = set()


def _composed(what: str, why: str) -> None:
    """The fused kernels have shape limits; outside them the op is COMPOSED from this package's kNN kernel and torch
    indexing (the reference's own op sequence) -- never silently: one warning per operator and reason."""
    key = what + why
    if key not in _WARNED:
        _WARNED.add(key)
        warnings.warn(f"pointcloudcounterfactual_b200.{what}: {why}; composing the op from knn + torch indexing "
                      "(slower than the fused sm_100a kernel)", RuntimeWarning, stacklevel=3)


def _fused_gather_ok(x: torch.Tensor, indices: torch.Tensor, k: int, what: str = "get_neighbours") -> bool:
    """The fused kernels cover CUDA fp32 features, int64 (B,N,k) indices, k <= 32, N <= 8192; anything else (and
    empty batches) takes the reference's torch composition, with a warning."""
    if not x.is_cuda:
        raise RuntimeError(f"{what}: CUDA tensors only (no CPU fallback in pointcloudcounterfactual_b200)")
    if x.shape[0] == 0:
        return False
    if x.dtype != torch.float32 or indices.dtype != torch.int64:
        _composed(what, f"dtypes {x.dtype} / {indices.dtype} (fused kernel: float32 features, int64 indices)")
        return False
    if indices.dim() != 3 or tuple(indices.shape) != (x.shape[0], x.shape[2], k):
        _composed(what, f"indices of shape {tuple(indices.shape)} (fused kernel: (B,N,k))")
        return False
    if k > 32 or x.shape[2] > 8192:
        _composed(what, f"k={k}, N={x.shape[2]} (fused kernel: k <= 32, N <= 8192)")
        return False
    return True


def get_neighbours(x: torch.Tensor, indices: torch.Tensor, k: int):
    """(:85-94) -> (indices (B,N,k), neighbours (B,C,N,k))."""
    batch, n_feat, n_points = x.size()
    if not indices.numel():
        indices = knn(x, k)
    if _fused_gather_ok(x, indices, k):
        return indices, _GraphGather.apply(x.contiguous(), indices.contiguous(), 0)
    flat = indices.contiguous().view(batch, 1, k * n_points).expand(-1, n_feat, -1)
    neighbours = torch.gather(x, 2, flat).view(batch, n_feat, n_points, k)
    return indices, neighbours


def get_local_covariance(x: torch.Tensor, indices: torch.Tensor, k: int = 16) -> torch.Tensor:
    """(:97-103) appends the flattened k-neighbourhood covariance to the features."""
    neighbours = get_neighbours(x, indices, k)[1]
    neighbours = neighbours - neighbours.mean(3, keepdim=True)
    cov = torch.matmul(neighbours.transpose(1, 2), neighbours.permute(0, 2, 3, 1))
    return torch.cat([x, cov.flatten(start_dim=2).transpose(1, 2)], dim=1).contiguous()


def graph_max_pooling(x: torch.Tensor, indices: torch.Tensor, k: int = 16) -> torch.Tensor:
    """(:106-110) max over the k neighbours of every point, (B,C,N) -> (B,C,N) (LDGCNN, src/module/encoders.py:84).
    On CUDA this is the fused EdgeConv edge pass with u_j = x_j, v_i = 0 and no normalisation (the operands are copied,
    not multiplied by an identity: no GEMM, so TF32 settings cannot touch the values): no (B,C,N,k) tensor,
    bit-identical values, the gradient goes to the first arg-max slot."""
    if not indices.numel():
        indices = knn(x, k)
    c = x.shape[1]
    if _fused_gather_ok(x, indices, k, "graph_max_pooling"):
        if c % 4 == 0 and 4 <= c <= 1024:
            from . import edgeconv  # late import: edgeconv builds on this module

            return edgeconv.graph_max_pool(x, indices)
        _composed("graph_max_pooling", f"C={c} (fused kernel: C % 4 == 0, 4 <= C <= 1024)")
    return get_neighbours(x, indices, k)[1].max(dim=-1)[0]


def get_graph_features(x: torch.Tensor, indices: torch.Tensor, k: int = 20) -> tuple[torch.Tensor, torch.Tensor]:
    """(:113-119) EdgeConv input: cat(neighbour - centre, centre) -> (B, 2C, N, k)."""
    if not indices.numel():
        indices = knn(x, k)
    if _fused_gather_ok(x, indices, k, "get_graph_features"):
        return indices, _GraphGather.apply(x.contiguous(), indices.contiguous(), 1)
    indices_out, neighbours = get_neighbours(x, indices, k)
    centre = x.unsqueeze(3).expand(-1, -1, -1, k)
    return indices_out, torch.cat([neighbours - centre, centre], dim=1).contiguous()


class _GraphFiltering(Function):
    """x (B,3,N), idx (B,N,k) -> smoothed cloud (B,3,N): one launch forward, one backward (pcc_graph_filtering*)."""

    @staticmethod
    def forward(ctx: Any, x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        b, _, n = x.shape
        k = idx.shape[2]
        with torch.cuda.device(x.device):
            out = torch.empty_like(x)
            mean = torch.empty((b,), dtype=torch.float32, device=x.device)
            L.check(L.load().pcc_graph_filtering(b, n, k, L.ptr(x), L.ptr(idx), L.ptr(out), L.ptr(mean), L.stream_of(x)),
                    "graph_filtering")
        ctx.save_for_backward(x, idx, mean)
        return out

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor):
        x, idx, mean = ctx.saved_tensors
        b, _, n = x.shape
        g = grad_out.contiguous()
        with torch.cuda.device(x.device):
            gx = torch.empty_like(x)
            L.check(L.load().pcc_graph_filtering_grad(b, n, idx.shape[2], L.ptr(x), L.ptr(idx), L.ptr(mean), L.ptr(g),
                                                      L.ptr(gx), L.stream_of(x)), "graph_filtering_grad")
        return gx, None


def graph_filtering(x: torch.Tensor, k: int = 4) -> torch.Tensor:
    """(:122-133) decoder-output smoothing; relies on ascending kNN with the point itself in column 0."""
    if (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and x.shape[1] == 3 and x.shape[0] > 0
            and 2 <= k <= 8 and k <= x.shape[2] <= 6144):
        xc = x.contiguous()
        return _GraphFiltering.apply(xc, knn(xc.detach(), k))
    if not x.is_cuda:
        raise RuntimeError("graph_filtering: CUDA tensors only (no CPU fallback in pointcloudcounterfactual_b200)")
    if x.shape[0] > 0:
        _composed("graph_filtering", f"input {tuple(x.shape)} {x.dtype}, k={k} (fused kernel: float32 (B,3,N), 2 <= k <= 8, "
                                     "k <= N <= 6144)")
    neighbours = get_neighbours(x, indices=torch.empty(0), k=k)[1][..., 1:]
    diff = x.unsqueeze(-1) - neighbours
    dist = torch.sqrt((diff * diff).sum(1).abs())
    sigma = torch.clamp(dist[..., 0:1].mean(1, keepdim=True), min=0.005)
    weights = torch.exp(-(dist / sigma))
    x_weight = weights.sum(2).unsqueeze(1)
    return (1 + x_weight) * x - (weights.unsqueeze(1) * neighbours).sum(-1)
