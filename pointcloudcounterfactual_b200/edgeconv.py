"""Fused EdgeConv layer: ``get_graph_features`` -> ``EdgeConvLayer`` (1x1 Conv2d, BatchNorm2d, activation) -> max
over the k neighbours, computed without the (B,2C,N,k) edge tensor.

Reference sequence (``src/module/encoders.py:49-54``, ``classifier.py:55-60``)::

    indices, x = get_graph_features(x, k=k, indices=indices)   # neighbour_ops.py:113-119 -> (B,2C,N,k)
    x = conv(x)                                                # layers.py:159-203: act(bn(dense(x)))
    x = x.max(dim=3, keepdim=False)[0]                         # (B,Cout,N)

Here the convolution weight W = [W1 | W2] is applied to the POINTS once -- ``u = W1 x``, ``v = (W2 - W1) x``, one
hand-written tcgen05 GEMM (3xTF32, fp32-accurate), 1/k of the reference's convolution work -- and ``pcc_edgeconv_forward`` does everything that touches
edges: y(i,t) = u[idx[i,t]] + v[i], the batch statistics of BatchNorm2d over all B*N*k edges, the per-channel affine
map, the activation and the max over k.  ``pcc_edgeconv_backward`` propagates through all of it (every edge gets a
gradient through the batch statistics), deterministically.  Results agree with the torch composition to fp32
rounding; the summation order inside the convolution differs (u + v instead of W . [x_j - x_i; x_i]).

There is no CPU fallback: CPU tensors raise.  Layers outside the fused kernel's limits (grouped or residual layers,
activations other than none / ReLU / LeakyReLU, k > 64, N > 8192, Cout % 4 != 0) take the reference's op sequence on
top of the fused gather.
"""
from __future__ import annotations

from typing import Any

import torch
from torch import nn
from torch.autograd import Function

from . import _lib as L
from .neighbour_ops import get_graph_features, knn

BN_EVAL, BN_TRAIN, AFFINE = 0, 1, 2
ACT_NONE, ACT_LEAKY = 0, 1


def gemm_nt(a: torch.Tensor, a_strides: tuple[int, int, int], b: torch.Tensor, b_strides: tuple[int, int, int],
            out: torch.Tensor, out_strides: tuple[int, int, int], batch: int, m: int, n: int, k: int,
            ksplit: int = 1) -> torch.Tensor:
    """out[z](i,j) = sum_l a[z](i,l) b[z](j,l) on the tcgen05 tensor cores with fp32-level accuracy (``pcc_gemm_tf32x3``,
    csrc/gemm_tc.cu); every operand is addressed through (batch, row, k) element strides, so neither x nor the gradient is
    ever transposed in memory.  ``ksplit`` > 1 writes batch * ksplit partial products (slice z * ksplit + p) for the caller
    to add: a long reduction spread over more CTAs."""
    L.check(L.load().pcc_gemm_tf32x3(batch, ksplit, m, n, k, L.ptr(a), *a_strides, L.ptr(b), *b_strides, L.ptr(out), *out_strides,
                                     L.stream_of(out)), "gemm_tf32x3")
    return out


class _EdgeConvMax(Function):
    """x (B,C,N) channels-first, weight (Cout,2C), idx (B,N,k) int64 -> out (B,Cout,N).

    The three GEMMs around the edge kernels run on the tcgen05 tensor cores with fp32-level accuracy (``gemm_nt`` ->
    ``pcc_gemm_tf32x3``: 3xTF32, operands split and laid out by the kernel's loader warps), written out here instead of
    left to autograd so that the weight gradient is a BATCHED product over clouds (K = N per problem) rather than one
    GEMM with K = B*N and two output tiles, and so that no transposed copy of x or of the gradient is made."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)  # the kernels are fp32: no autocast inside
    def forward(ctx: Any, x: torch.Tensor, weight: torch.Tensor, idx: torch.Tensor, gamma: torch.Tensor | None,
                beta: torch.Tensor | None, running_mean: torch.Tensor | None, running_var: torch.Tensor | None,
                bn_mode: int, momentum: float, eps: float, act: int, slope: float) -> torch.Tensor:
        b, c, n = x.shape
        cout = weight.shape[0]
        k = idx.shape[2]
        dev = x.device
        w1 = weight[:, :c]
        ws = torch.cat([w1, weight[:, c:] - w1], dim=0)            # (2Cout, C): [W1 ; W2 - W1]
        with torch.cuda.device(dev):
            # uv[b,i,:] = [W1 x_i | (W2-W1) x_i], point-major rows, straight from the channels-first x (any strides)
            uv = torch.empty((b, n, 2 * cout), dtype=torch.float32, device=dev)
            ws = ws.contiguous()
            gemm_nt(x, (x.stride(0), x.stride(2), x.stride(1)), ws, (0, c, 1), uv, (n * 2 * cout, 2 * cout, 1),
                    b, n, 2 * cout, c)
            out = torch.empty((b, cout, n), dtype=torch.float32, device=dev)
            exty = torch.empty((b, n, cout), dtype=torch.float32, device=dev)
            sy = torch.empty((b, n, cout), dtype=torch.float32, device=dev) if bn_mode == BN_TRAIN else None
            slot = torch.empty((b, (cout + 7) // 8, n, 8), dtype=torch.uint8, device=dev)  # slice-major, see edgeconv.cu
            mean = torch.empty((cout,), dtype=torch.float32, device=dev)
            invstd = torch.empty((cout,), dtype=torch.float32, device=dev)
            L.check(L.load().pcc_edgeconv_forward(
                b, n, k, cout, L.ptr(uv), L.ptr(idx), L.ptr(gamma), L.ptr(beta), L.ptr(running_mean),
                L.ptr(running_var), bn_mode, momentum, eps, act, slope, L.ptr(out), L.ptr(exty), L.ptr(sy),
                L.ptr(slot), L.ptr(mean), L.ptr(invstd), L.stream_of(uv)), "edgeconv_forward")
        ctx.save_for_backward(x, ws, uv, idx, gamma, beta, mean, invstd, exty, sy, slot)
        ctx.cfg = (bn_mode, act, slope)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx: Any, grad_out: torch.Tensor):
        x, ws, uv, idx, gamma, beta, mean, invstd, exty, sy, slot = ctx.saved_tensors
        bn_mode, act, slope = ctx.cfg
        b, n, c2 = uv.shape
        cout = c2 // 2
        c = x.shape[1]
        g = grad_out.contiguous()
        dev = uv.device
        with torch.cuda.device(dev):
            guv = torch.empty_like(uv)
            ggamma = torch.empty((cout,), dtype=torch.float32, device=dev) if gamma is not None else None
            gbeta = torch.empty((cout,), dtype=torch.float32, device=dev) if beta is not None else None
            L.check(L.load().pcc_edgeconv_backward(
                b, n, idx.shape[2], cout, L.ptr(uv), L.ptr(idx), L.ptr(gamma), L.ptr(beta), L.ptr(mean), L.ptr(invstd),
                bn_mode, act, slope, L.ptr(exty), L.ptr(sy), L.ptr(slot), L.ptr(g), L.ptr(guv), L.ptr(ggamma),
                L.ptr(gbeta), L.stream_of(uv)), "edgeconv_backward")
            gx = gw = None
            if ctx.needs_input_grad[0]:   # gx^T (N,C) = guv (N,2Cout) . ws (2Cout,C), stored channels-first
                gx = torch.empty((b, c, n), dtype=torch.float32, device=dev)
                gemm_nt(guv, (n * c2, c2, 1), ws, (0, 1, c), gx, (c * n, 1, n), b, n, c, c2)
            if ctx.needs_input_grad[1]:   # per cloud (2Cout,N) . (N,C), then summed over the clouds
                ksplit = max(1, min(8, n // 512))  # K = N points per cloud: a few hundred CTAs instead of b * ceil(2Cout/128)
                gwb = torch.empty((b * ksplit, c2, c), dtype=torch.float32, device=dev)
                gemm_nt(guv, (n * c2, 1, c2), x, (x.stride(0), x.stride(1), x.stride(2)), gwb, (c2 * c, c, 1), b, c2, c, n,
                        ksplit)
                gws = gwb.sum(0)
                gw = torch.cat([gws[:cout] - gws[cout:], gws[cout:]], dim=1)
        return (gx, gw, None, ggamma if ctx.needs_input_grad[3] else None, gbeta if ctx.needs_input_grad[4] else None,
                None, None, None, None, None, None, None)


class _GraphMaxPool(Function):
    """x (B,C,N), idx (B,N,k) -> max over the k neighbours (B,C,N): the edge pass with u = x, v = 0 and no
    normalisation.  uv = [x^T | 0] is built by a copy -- no GEMM, so the values are bit-identical to gather + max whatever
    ``torch.backends.cuda.matmul.allow_tf32`` says; the gradient goes to the first arg-max slot."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx: Any, x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        b, c, n = x.shape
        k = idx.shape[2]
        dev = x.device
        with torch.cuda.device(dev):
            uv = torch.zeros((b, n, 2 * c), dtype=torch.float32, device=dev)
            uv[:, :, :c].copy_(x.transpose(1, 2))
            out = torch.empty((b, c, n), dtype=torch.float32, device=dev)
            exty = torch.empty((b, n, c), dtype=torch.float32, device=dev)
            slot = torch.empty((b, (c + 7) // 8, n, 8), dtype=torch.uint8, device=dev)
            mean = torch.empty((c,), dtype=torch.float32, device=dev)
            invstd = torch.empty((c,), dtype=torch.float32, device=dev)
            L.check(L.load().pcc_edgeconv_forward(
                b, n, k, c, L.ptr(uv), L.ptr(idx), None, None, None, None, AFFINE, 0.0, 0.0, ACT_NONE, 0.0, L.ptr(out),
                L.ptr(exty), None, L.ptr(slot), L.ptr(mean), L.ptr(invstd), L.stream_of(uv)), "graph_max_pooling")
        ctx.save_for_backward(uv, idx, mean, invstd, exty, slot)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx: Any, grad_out: torch.Tensor):
        uv, idx, mean, invstd, exty, slot = ctx.saved_tensors
        b, n, c2 = uv.shape
        c = c2 // 2
        g = grad_out.contiguous()
        with torch.cuda.device(uv.device):
            guv = torch.empty_like(uv)
            L.check(L.load().pcc_edgeconv_backward(
                b, n, idx.shape[2], c, L.ptr(uv), L.ptr(idx), None, None, L.ptr(mean), L.ptr(invstd), AFFINE, ACT_NONE,
                0.0, L.ptr(exty), None, L.ptr(slot), L.ptr(g), L.ptr(guv), None, None, L.stream_of(uv)),
                "graph_max_pooling backward")
        return guv[:, :, :c].transpose(1, 2).contiguous(), None


def graph_max_pool(x: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    L.require_cuda(x, contiguous=False)
    return _GraphMaxPool.apply(x, indices.contiguous())


def _act_code(act: nn.Module | None) -> tuple[int, float] | None:
    if act is None or isinstance(act, nn.Identity):
        return ACT_NONE, 0.0
    if isinstance(act, nn.LeakyReLU) and act.negative_slope >= 0:
        return ACT_LEAKY, float(act.negative_slope)
    if isinstance(act, nn.ReLU):
        return ACT_LEAKY, 0.0
    return None


def fused_limits_ok(x: torch.Tensor, indices: torch.Tensor, k: int, cout: int) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and x.shape[0] > 0 and indices.dtype == torch.int64
            and tuple(indices.shape) == (x.shape[0], x.shape[2], k) and k <= 64 and x.shape[2] <= 8192
            and cout % 4 == 0 and 4 <= cout <= 1024)


def edge_conv_max(x: torch.Tensor, indices: torch.Tensor, weight: torch.Tensor, bn_weight: torch.Tensor | None = None,
                  bn_bias: torch.Tensor | None = None, running_mean: torch.Tensor | None = None,
                  running_var: torch.Tensor | None = None, bn_mode: int = BN_TRAIN, momentum: float = 0.1,
                  eps: float = 1e-5, negative_slope: float | None = None) -> torch.Tensor:
    """x (B,C,N) channels-first, indices (B,N,k) int64, weight (Cout,2C[,1,1]) -> (B,Cout,N).

    Equals ``act(batch_norm(conv2d(graph_features(x, indices), weight))).max(3)[0]`` with ``act`` the identity
    (``negative_slope=None``) or LeakyReLU / ReLU (``negative_slope=0``).  ``bn_mode``: BN_TRAIN batch statistics
    (running stats updated in place), BN_EVAL running statistics, AFFINE no normalisation (``bn_bias`` = conv bias).
    """
    L.require_cuda(x, contiguous=False)
    cout = weight.shape[0]
    act, slope = (ACT_NONE, 0.0) if negative_slope is None else (ACT_LEAKY, float(negative_slope))
    return _EdgeConvMax.apply(x, weight.reshape(cout, 2 * x.shape[1]), indices.contiguous(), bn_weight, bn_bias,
                              running_mean, running_var, bn_mode, float(momentum), float(eps), act, slope)


def fused_edge_conv(layer: nn.Module, x: torch.Tensor, indices: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Drop-in for the three reference lines quoted in the module docstring, for an ``EdgeConvLayer``-shaped module
    (attributes ``dense`` = Conv2d 1x1, ``bn`` = BatchNorm2d or None, ``act`` = module or None, ``residual``):
    returns ``(indices, max_k act(bn(dense(graph_features))))`` with x (B,C,N) and the result (B,Cout,N)."""
    if not indices.numel():
        indices = knn(x, k)
    dense, bn, act = layer.dense, getattr(layer, "bn", None), getattr(layer, "act", None)
    code = _act_code(act)
    ok = (code is not None and isinstance(dense, nn.Conv2d) and dense.kernel_size == (1, 1) and dense.groups == 1
          and not getattr(layer, "residual", False) and (bn is None or isinstance(bn, nn.BatchNorm2d))
          and fused_limits_ok(x, indices, k, dense.out_channels))
    if ok and bn is not None:
        batch_stats = bn.training or bn.running_mean is None
        ok = bn.momentum is not None or not batch_stats  # cumulative moving average is left to torch
    if not ok:
        feat = get_graph_features(x, indices, k)[1]
        return indices, layer(feat).max(dim=3, keepdim=False)[0]
    slope = None if code[0] == ACT_NONE else code[1]
    if bn is None:
        out = edge_conv_max(x, indices, dense.weight, None, dense.bias, bn_mode=AFFINE, negative_slope=slope)
        return indices, out
    if batch_stats and bn.training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    out = edge_conv_max(x, indices, dense.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                        BN_TRAIN if batch_stats else BN_EVAL, bn.momentum if bn.momentum is not None else 0.1, bn.eps,
                        slope)
    return indices, out


def dgcnn_forward(edge_convolutions, final_conv: nn.Module, x: torch.Tensor, indices: torch.Tensor,
                  n_neighbors: int) -> torch.Tensor:
    """``DGCNN.forward`` (src/module/encoders.py:45-59) on the fused layer: x (B,N,3) -> (B,w_dim)."""
    xs = []
    x = x.transpose(2, 1)
    for conv in edge_convolutions:
        indices, x = fused_edge_conv(conv, x, indices, n_neighbors)
        indices = torch.empty(0)  # neighbours are recomputed in feature space every layer
        xs.append(x)
    x = torch.cat(xs, dim=1).contiguous()
    return final_conv(x).max(dim=2, keepdim=False)[0]
