"""Drop-in for the reference's pybind module ``emd_backend`` (external/emd/src/emd.cpp:28-31): ``forward`` with the
same 16 positional arguments and ``backward`` with the same 5, returning the same int codes (1 ok, -1 shape error).
Unlike the reference the work is enqueued on the CURRENT stream of the tensors' device."""
from __future__ import annotations

import torch

from .. import _lib as L


def forward(xyz1, xyz2, dist, assignment, price, assignment_inv, bid, bid_increments, max_increments, unass_idx,
            unass_cnt, unass_cnt_sum, cnt_tmp, max_idx, eps: float, iters: int) -> int:
    L.require_cuda(xyz1, xyz2, dist, price, bid_increments, max_increments)
    L.require_cuda(assignment, assignment_inv, bid, unass_idx, unass_cnt, unass_cnt_sum, cnt_tmp, max_idx,
                   dtype=torch.int32)
    b, n, m = xyz1.size(0), xyz1.size(1), xyz2.size(1)
    with torch.cuda.device(xyz1.device):
        rc = L.load().pcc_emd_forward(b, n, m, L.ptr(xyz1), L.ptr(xyz2), L.ptr(dist), L.ptr(assignment), L.ptr(price),
                                      L.ptr(assignment_inv), L.ptr(bid), L.ptr(bid_increments), L.ptr(max_increments),
                                      L.ptr(unass_idx), L.ptr(unass_cnt), L.ptr(unass_cnt_sum), L.ptr(cnt_tmp),
                                      L.ptr(max_idx), float(eps), int(iters), L.stream_of(xyz1))
    if rc not in (1, -1):
        L.check(rc, "emd_backend.forward")
    return rc


def backward(xyz1, xyz2, gradxyz, graddist, idx) -> int:
    L.require_cuda(xyz1, xyz2, gradxyz, graddist)
    L.require_cuda(idx, dtype=torch.int32)
    with torch.cuda.device(xyz1.device):
        rc = L.load().pcc_emd_backward(xyz1.size(0), xyz1.size(1), L.ptr(xyz1), L.ptr(xyz2), L.ptr(gradxyz),
                                       L.ptr(graddist), L.ptr(idx), L.stream_of(xyz1))
    if rc != 1:
        L.check(rc, "emd_backend.backward")
    return rc
