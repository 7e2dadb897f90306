"""``emdModule`` (auction-algorithm EMD), mirroring the reference's external/emd/emd/emd_module.py:16-100.

Same call ``emdModule()(input1, input2, eps, iters) -> (dist, assignment)``, same shape rules and ValueErrors,
gradient only w.r.t. the first cloud.  Work buffers are allocated on the inputs' device (the reference hard-codes
``device='cuda'``, emd_module.py:34-45) and the kernel runs on the current stream.
"""
from __future__ import annotations

from typing import Any

import torch
from torch import nn
from torch.autograd import Function

from . import emd_backend


class emdFunction(Function):
    @staticmethod
    def forward(ctx: Any, xyz1: torch.Tensor, xyz2: torch.Tensor, eps: float, iters: int):
        batch1, n, _ = xyz1.size()
        batch2, m, _ = xyz2.size()
        if n != m:
            raise ValueError('Input point clouds should have the same number of points')
        if batch1 != batch2:
            raise ValueError('Batch size must be the same')
        if n % 1024:
            raise ValueError('Only valid for clouds of a size multiple of 1024')
        if batch1 > 512:
            raise ValueError('Batch size should not exceed 512')
        if not xyz1.is_cuda:
            xyz1 = xyz1.cuda()
        dev = xyz1.device
        xyz1 = xyz1.contiguous().float()
        xyz2 = xyz2.to(dev).contiguous().float()

        def buf(shape, dtype, fill=0):
            return torch.full(shape, fill, dtype=dtype, device=dev)

        dist = buf((batch1, n), torch.float32)
        assignment = buf((batch1, n), torch.int32, -1)
        assignment_inv = buf((batch1, m), torch.int32, -1)
        price = buf((batch1, m), torch.float32)
        bid = buf((batch1, n), torch.int32)
        bid_increments = buf((batch1, n), torch.float32)
        max_increments = buf((batch1, m), torch.float32)
        unass_idx = buf((batch1 * n,), torch.int32)
        max_idx = buf((batch1 * m,), torch.int32)
        unass_cnt = buf((512,), torch.int32)
        unass_cnt_sum = buf((512,), torch.int32)
        cnt_tmp = buf((512,), torch.int32)
        emd_backend.forward(xyz1, xyz2, dist, assignment, price, assignment_inv, bid, bid_increments, max_increments,
                            unass_idx, unass_cnt, unass_cnt_sum, cnt_tmp, max_idx, eps, iters)
        ctx.save_for_backward(xyz1, xyz2, assignment)
        ctx.mark_non_differentiable(assignment)
        return dist, assignment

    @staticmethod
    def backward(ctx: Any, grad_dist: torch.Tensor, _grad_assignment: Any = None):
        xyz1, xyz2, assignment = ctx.saved_tensors
        grad_xyz1 = torch.empty_like(xyz1)
        emd_backend.backward(xyz1, xyz2, grad_xyz1, grad_dist.contiguous(), assignment)
        return grad_xyz1, torch.zeros_like(xyz2), None, None


class emdModule(nn.Module):
    """``forward(input1, input2, eps, iters) -> (dist (B,N) squared distance to the match, assignment (B,N) int32)``
    for clouds normalised to [0, 1], N a multiple of 1024, B <= 512 (external/emd/README.md:17)."""

    def forward(self, input1: torch.Tensor, input2: torch.Tensor, eps: float, iters: int):
        return emdFunction.apply(input1, input2, eps, iters)
