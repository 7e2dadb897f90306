"""Batch sharding of the geometry hot path over ranks (one process per GPU) and loss / gradient reduction.

The hot-path operators are independent per cloud (every reference kernel loops ``for (i = blockIdx.x; i < b; ...)``,
e.g. nndistance.cu:5, approxmatch.cu:15), so multi-GPU execution is a contiguous split of the batch -- the
reference's own data-parallel contract (``src/utils/parallel.py:42-53``: one process per GPU,
``src/config/specs.py:339-345``: batch_size // n_subprocesses per rank) -- with NO collective inside the operators.
The only exchanges are an all-reduce of the loss for reporting and, when training, of the gradients (NCCL over
NVLink 5 / NVSwitch on a B200 box; gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Callable, Sequence

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of a global batch owned by ``rank``; sizes differ by at most one cloud."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("invalid rank / world_size")
    base, extra = divmod(batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors: Sequence[torch.Tensor], world_size: int, rank: int) -> list[torch.Tensor]:
    """Slice every (B, ...) tensor to this rank's clouds (views, no copy)."""
    out = []
    for t in tensors:
        lo, hi = shard_bounds(t.size(0), world_size, rank)
        out.append(t[lo:hi])
    return out


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Join the process group described by RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).
    Returns (rank, world_size, local_rank); a single-process run needs no process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def global_mean_loss(per_cloud: torch.Tensor) -> torch.Tensor:
    """Mean of a per-cloud loss over ALL ranks' clouds: one all-reduce of (sum, count)."""
    acc = torch.stack([per_cloud.detach().sum().to(torch.float64),
                       torch.tensor(float(per_cloud.numel()), dtype=torch.float64, device=per_cloud.device)])
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return (acc[0] / acc[1].clamp(min=1.0)).to(per_cloud.dtype)


class PendingMean:
    """Global mean whose all-reduce is in flight (``global_mean_loss_async``).  ``wait()`` makes the current stream wait
    for the collective and returns the mean; until then the launching stream keeps running -- a reporting collective has
    no business on the critical path of the step that produced it."""

    def __init__(self, acc: torch.Tensor, work, dtype: torch.dtype):
        self._acc, self._work, self._dtype = acc, work, dtype

    def wait(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        return (self._acc[0] / self._acc[1].clamp(min=1.0)).to(self._dtype)


def global_mean_loss_async(per_cloud: torch.Tensor) -> PendingMean:
    """``global_mean_loss`` with the all-reduce launched asynchronously (NCCL: on the communicator's own stream)."""
    acc = torch.stack([per_cloud.detach().sum().to(torch.float64),
                       torch.tensor(float(per_cloud.numel()), dtype=torch.float64, device=per_cloud.device)])
    work = None
    if dist.is_initialized() and dist.get_world_size() > 1:
        work = dist.all_reduce(acc, op=dist.ReduceOp.SUM, async_op=True)
    return PendingMean(acc, work, per_cloud.dtype)


class LossAccumulator:
    """Running (sum, count) of per-cloud losses ON THE DEVICE, reduced over the ranks once per logging interval.

    The reference never exchanges the loss per step: drytorch aggregates a metric locally and synchronises it when the
    epoch's value is read.  A per-step all-reduce of 16 bytes costs nothing in bandwidth but couples the ranks -- every
    step then ends when the SLOWEST rank's previous step has ended (measured: 4-7 % of weak-scaling efficiency on 8
    GPUs).  ``add`` is one fused elementwise update with no communication; ``reduce`` is ONE all-reduce of two doubles."""

    def __init__(self, device: torch.device):
        self._acc = torch.zeros(2, dtype=torch.float64, device=device)

    def add(self, per_cloud: torch.Tensor) -> None:
        self._acc[0] += per_cloud.detach().sum(dtype=torch.float64)
        self._acc[1] += float(per_cloud.numel())

    def reduce(self, reset: bool = True) -> torch.Tensor:
        """Global mean over every cloud added on every rank since the last reset (a 0-dim float64 device tensor)."""
        acc = self._acc.clone()
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        if reset:
            self._acc.zero_()
        return acc[0] / acc[1].clamp(min=1.0)


def all_reduce_mean_(grad: torch.Tensor) -> torch.Tensor:
    """In-place data-parallel gradient averaging of a flat buffer (what DDP does per bucket)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM)
        grad.div_(dist.get_world_size())
    return grad


class ShardedLoss:
    """Runs ``loss_fn(recon_shard, ref_shard) -> (b_local,)`` on this rank's slice of a global batch.

    ``__call__`` returns (per-cloud losses of the local shard, global mean over all ranks).  ``loss_fn`` is the
    operator under test (``losses.chamfer_emd`` on GPUs; any callable in the gloo CPU tests).
    """

    def __init__(self, loss_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], rank: int, world_size: int):
        self.loss_fn, self.rank, self.world_size = loss_fn, rank, world_size

    def __call__(self, recon: torch.Tensor, ref: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        r, t = shard_batch([recon, ref], self.world_size, self.rank)
        local = self.loss_fn(r, t) if r.size(0) else r.new_zeros((0,))
        return local, global_mean_loss(local)
