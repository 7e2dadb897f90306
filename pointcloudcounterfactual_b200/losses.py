"""Chamfer reconstruction losses with the reference's names and reductions
(``src/train/metrics_and_losses.py:21-47``), computed by the sm_100a ``nn_distance`` operator.

The reference's two back-ends disagree on the reduction: ``pykeops_chamfer`` (the one
used in GPU training) takes the MEAN over points, ``torch_chamfer`` the SUM.  Both are kept as they are.
"""
from __future__ import annotations

from typing import Any

import torch
from torch.autograd import Function

from . import _lib as L
from .structural_losses import match_cost, nn_distance


class _ChamferReduce(Function):
    """loss[b] = s1 * sum_j min_k |t1_j - t2_k|^2 + s2 * sum_k min_j |t2_k - t1_j|^2 in two launches + one tiny
    reduction, backward in ONE launch (pcc_chamfer_reduce / pcc_chamfer_reduce_grad): the same nearest-neighbour pairs
    and gradients as ``nn_distance`` followed by torch reductions, without the per-point gradient arrays."""

    @staticmethod
    def forward(ctx: Any, t1: torch.Tensor, t2: torch.Tensor, mean: bool) -> torch.Tensor:
        L.require_cuda(t1, t2)
        if t1.dim() != 3 or t2.dim() != 3 or t1.size(2) != 3 or t2.size(2) != 3 or t1.size(0) != t2.size(0):
            raise RuntimeError("expected point sets of shape (batch, points, 3) with equal batch sizes")
        b, n, m = t1.size(0), t1.size(1), t2.size(1)
        s1, s2 = (1.0 / n, 1.0 / m) if mean else (1.0, 1.0)
        with torch.cuda.device(t1.device):
            loss = torch.empty((b,), dtype=torch.float32, device=t1.device)
            d1 = torch.empty((b, n), dtype=torch.float32, device=t1.device)
            d2 = torch.empty((b, m), dtype=torch.float32, device=t1.device)
            i1 = torch.empty((b, n), dtype=torch.int32, device=t1.device)
            i2 = torch.empty((b, m), dtype=torch.int32, device=t1.device)
            L.check(L.load().pcc_chamfer_reduce(b, n, L.ptr(t1), m, L.ptr(t2), s1, s2, L.ptr(loss), L.ptr(d1), L.ptr(i1),
                                                L.ptr(d2), L.ptr(i2), L.stream_of(t1)), "chamfer_reduce")
        ctx.save_for_backward(t1, t2, i1, i2)
        ctx.scales = (s1, s2)
        return loss

    @staticmethod
    def backward(ctx: Any, grad_loss: torch.Tensor):
        t1, t2, i1, i2 = ctx.saved_tensors
        s1, s2 = ctx.scales
        b, n, m = t1.size(0), t1.size(1), t2.size(1)
        g = grad_loss.contiguous().float()
        with torch.cuda.device(t1.device):
            g1 = torch.empty_like(t1)
            g2 = torch.empty_like(t2)
            rc = L.load().pcc_chamfer_reduce_grad(b, n, L.ptr(t1), m, L.ptr(t2), L.ptr(i1), L.ptr(i2), L.ptr(g), s1, s2,
                                                  L.ptr(g1), L.ptr(g2), L.stream_of(t1))
        if rc == -2:  # PCC_ENOTSUP (clouds above 16k points): per-point upstream gradients through NNDistanceGrad
            from .structural_losses.structural_losses_backend import NNDistanceGrad
            g1, g2 = NNDistanceGrad(t1, t2, i1, i2, (g * s1)[:, None].expand(b, n).contiguous(),
                                    (g * s2)[:, None].expand(b, m).contiguous())
        else:
            L.check(rc, "chamfer_reduce_grad")
        return g1, g2, None


def _chamfer(t1: torch.Tensor, t2: torch.Tensor, mean: bool) -> torch.Tensor:
    t1, t2 = t1.contiguous(), t2.contiguous()
    if t1.size(1) == 0 or t2.size(1) == 0 or t1.size(0) == 0:  # degenerate shapes: the unfused composition
        dist1, dist2 = nn_distance(t1, t2)
        return dist2.mean(1) + dist1.mean(1) if mean else dist1.sum(1) + dist2.sum(1)
    return _ChamferReduce.apply(t1, t2, mean)


def pykeops_chamfer(t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor:
    """(:21-41) (B,N,3),(B,M,3) -> (B,): mean_j min_k |t1_j - t2_k|^2 + mean_k min_j |t2_k - t1_j|^2.
    Gradients reach both clouds through the nearest-neighbour pairs, like the reference's gather-based form."""
    return _chamfer(t1, t2, True)


def torch_chamfer(t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor:
    """(:44-47) same pairs, SUM over points."""
    return _chamfer(t1, t2, False)


def chamfer_emd(recon: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """The reference's ChamferEMD reconstruction loss on GPU (:70-79 -> :59-67 + :50-56):
    ``pykeops_chamfer(recon, ref) + match_cost(recon, ref)``, one value per cloud."""
    return pykeops_chamfer(recon, ref) + match_cost(recon.contiguous(), ref.contiguous())


class GraphedLossStep:
    """One training-loss step for FIXED shapes as a CUDA graph: [host clouds -> device,] ``loss_fn(recon, ref)`` forward,
    backward w.r.t. ``recon``, per-cloud loss back to the host.  Replaying it costs one launch instead of ~60 Python-side
    calls, so the GPU does not idle between kernels or between steps while the host issues work (B=32 x 2048, end to
    end: 16.8 k -> 18.4 k clouds/s).  Usage::

        step = GraphedLossStep(chamfer_emd, recon_host, ref_host, device)   # pinned (B,N,3) / (B,M,3) host tensors
        recon_host.copy_(new_batch)            # refill the SAME pinned buffers, then
        loss_host, grad = step()               # loss_host: pinned (B,), valid after step.synchronize(); grad: device (B,N,3)

    With CUDA tensors instead of pinned host tensors the copies are left out and ``step.recon`` / ``step.ref`` are the
    device buffers to refill.  ``loss_device`` holds the per-cloud loss on the device (e.g. for
    ``sharding.global_mean_loss_async``)."""

    def __init__(self, loss_fn, recon: torch.Tensor, ref: torch.Tensor, device: torch.device, warmup: int = 2,
                 staged: bool = False):
        self.device = torch.device(device)
        self._staged = staged  # device inputs are staging buffers: the graph starts by copying them (PipelinedLossStep)
        self._from_host = not recon.is_cuda
        if self._from_host and not (recon.is_pinned() and ref.is_pinned()):
            raise RuntimeError("GraphedLossStep needs pinned host tensors (the copies are part of the graph)")
        self._recon_h, self._ref_h = recon, ref
        with torch.cuda.device(self.device):
            if self._from_host:
                self.recon = torch.empty(recon.shape, dtype=recon.dtype, device=self.device).requires_grad_(True)
                self.ref = torch.empty(ref.shape, dtype=ref.dtype, device=self.device)
            else:
                self.recon = recon.detach().clone().requires_grad_(True)
                self.ref = ref.detach().clone()
            self.loss_host = torch.empty((recon.size(0),), dtype=torch.float32).pin_memory()
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):  # warm-up outside the capture (lazy initialisation, memory pool)
                for _ in range(max(warmup, 1)):
                    self._body(loss_fn)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            self._graph = torch.cuda.CUDAGraph()
            n0 = L.launch_count()
            with torch.cuda.graph(self._graph):
                self._body(loss_fn)
            self.kernels_per_replay = L.launch_count() - n0  # libpcc_b200 kernels inside one replay

    def _body(self, loss_fn) -> None:
        if self._from_host or self._staged:
            with torch.no_grad():
                self.recon.copy_(self._recon_h, non_blocking=True)
                self.ref.copy_(self._ref_h, non_blocking=True)
        loss = loss_fn(self.recon, self.ref)
        (self.grad,) = torch.autograd.grad(loss.sum(), self.recon)
        self.loss_device = loss.detach()
        self.loss_host.copy_(self.loss_device, non_blocking=True)

    def __call__(self) -> tuple[torch.Tensor, torch.Tensor]:
        self._graph.replay()
        return self.loss_host, self.grad

    def synchronize(self) -> None:
        torch.cuda.current_stream(self.device).synchronize()


class PipelinedLossStep:
    """``GraphedLossStep`` with the host-to-device copies taken off the critical path: two device staging buffers and two
    captured graphs alternate, and while the graph of step i runs, the pinned host clouds of step i+1 travel to the
    other staging buffer on a copy stream.  Every step still moves its inputs host -> device and its per-cloud loss
    device -> host; only the waiting is overlapped.  Usage::

        pipe = PipelinedLossStep(chamfer_emd, recon_host, ref_host, device)   # pinned host tensors holding batch 0
        pipe.prefetch()                    # batch 0 starts travelling to the device
        for i in range(steps):
            pipe.wait_prefetch()           # the DMA that reads the pinned buffers has finished: only now may they be
            fill(recon_host, ref_host)     # rewritten, with batch i+1 (rewriting earlier races with the copy in flight)
            loss_prev = pipe.step()        # runs batch i, starts the copy of batch i+1, returns the per-cloud loss of
            ...                            # batch i-1 on the host (None on the first call)
        last = pipe.drain()
    """

    def __init__(self, loss_fn, recon_host: torch.Tensor, ref_host: torch.Tensor, device: torch.device):
        if not (recon_host.is_pinned() and ref_host.is_pinned()):
            raise RuntimeError("PipelinedLossStep needs pinned host tensors")
        self.device = torch.device(device)
        self._recon_h, self._ref_h = recon_host, ref_host
        with torch.cuda.device(self.device):
            self._copy = torch.cuda.Stream(self.device)
            self._stage = [(torch.empty(recon_host.shape, dtype=recon_host.dtype, device=self.device),
                            torch.empty(ref_host.shape, dtype=ref_host.dtype, device=self.device)) for _ in range(2)]
            self._steps = [GraphedLossStep(loss_fn, a, b, self.device, staged=True) for a, b in self._stage]
            self._ready = [torch.cuda.Event() for _ in range(2)]   # staging buffer filled
            self._done = [torch.cuda.Event() for _ in range(2)]    # graph finished (staging free, loss on the host)
        self._next = 0        # staging buffer the next prefetch fills
        self._pending = None  # index of the step whose loss has not been returned yet
        self.kernels_per_replay = self._steps[0].kernels_per_replay
        self.grad = None
        self.loss_device = None

    def prefetch(self) -> None:
        i = self._next
        self._copy.wait_event(self._done[i])  # the graph that read this staging buffer last has finished
        with torch.cuda.stream(self._copy):
            self._stage[i][0].copy_(self._recon_h, non_blocking=True)
            self._stage[i][1].copy_(self._ref_h, non_blocking=True)
            self._ready[i].record(self._copy)

    def wait_prefetch(self) -> None:
        """Block the host until the most recent ``prefetch()`` (the one ``step()`` issues included) has read the pinned
        host buffers -- the only moment from which they may be refilled.  Waits for that copy alone, not for the device."""
        self._ready[self._next].synchronize()

    def step(self):
        i = self._next
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._ready[i])
        _, self.grad = self._steps[i]()
        self.loss_device = self._steps[i].loss_device
        self._done[i].record(cur)
        self._next = 1 - i
        self.prefetch()  # the next step's inputs travel while this step computes
        prev = self._collect()
        self._pending = i
        return prev

    def _collect(self):
        if self._pending is None:
            return None
        self._done[self._pending].synchronize()
        return self._steps[self._pending].loss_host

    def drain(self):
        out = self._collect()
        self._pending = None
        return out
