"""EMD solver sweep (am_sweep_kernel) under the decompositions the library can launch: points per thread P = 1, 2, 4 through
the measurement hook pcc_approxmatch_sweep -- time per sweep and achieved Gexp/s against the MUFU.EX2 peak.
    python tools/emd_sweep_variants.py > profiles/r02_emd_sweep_variants.log"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import _lib, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
B, N = 32, 2048
lib = _lib.load()
a, c = (t.to(dev) for t in synthetic.s1_near(B, N))
ones = torch.ones(B, N, device=dev)
ratio = torch.empty(B, N, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
pairs = B * N * N
for P in (1, 2, 4):
    def sweep():
        _lib.check(lib.pcc_approxmatch_sweep(B, N, N, a.data_ptr(), c.data_ptr(), ones.data_ptr(), ones.data_ptr(),
                                             ratio.data_ptr(), -16.0, P, st), "sweep")
    try:
        for _ in range(5):
            sweep()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(100):
            sweep()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 10
        print(f"points per thread P={P}: {us:6.1f} us per sweep, {pairs / us / 1e3:7.1f} Gexp/s "
              f"({pairs / us / 1e3 / 4627 * 100:4.1f} % of the measured MUFU.EX2 peak 4627 Gexp/s); "
              f"warps = {B * N // (32 * P)} on 592 SM sub-partitions = {B * N / (32 * P) / 592:.2f} per sub-partition")
    except RuntimeError as ex:
        print(f"points per thread P={P}: not available ({ex})")
