"""Kernel tuning probes (run on the GPU box): approxmatch sweep variants, op timings under CUDA-graph replay."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import _lib, neighbour_ops, synthetic  # noqa: E402
from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import (  # noqa: E402
    ApproxMatch, MatchCostFused, NNDistance, NNDistanceGrad)

dev = torch.device("cuda", 0)
lib = _lib.load()
B, N = 32, 2048


def ev(fn, reps=30, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # microseconds


def graph(fn, reps=30):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return ev(g.replay, reps)


recon, ref = (t.to(dev) for t in synthetic.s1_near(B, N))
ones = torch.ones(B, N, device=dev)
ratio = torch.empty(B, N, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
out = {}
for level in (-16.0,):
    for p in (2, 4, 102, 202, 104):
        def sweep(p=p, level=level):
            _lib.check(lib.pcc_approxmatch_sweep(B, N, N, recon.data_ptr(), ref.data_ptr(), ones.data_ptr(), ones.data_ptr(),
                                                 ratio.data_ptr(), level, p, st), "sweep")
        out[f"sweep_us_P{p}_level{level}"] = ev(sweep)
out["nndistance_us"] = graph(lambda: NNDistance(recon, ref))
d1, i1, d2, i2 = NNDistance(recon, ref)
g1 = torch.ones_like(d1) / N
out["nndistancegrad_us"] = graph(lambda: NNDistanceGrad(recon, ref, i1, i2, g1, g1))
out["matchcost_fused_us"] = graph(lambda: MatchCostFused(recon, ref, True, False), reps=10)
out["approxmatch_us"] = ev(lambda: ApproxMatch(recon, ref), reps=5)
for name, x, k in (("knn_xyz_k20_n1024", synthetic.knn_xyz(B, 1024), 20), ("knn_feat64_k20_n1024", synthetic.knn_features(B, 64, 1024), 20),
                   ("knn_xyz_k4_n2048", synthetic.knn_xyz(B, 2048), 4)):
    xd = x.to(dev)
    out[name + "_us"] = graph(lambda xd=xd, k=k: neighbour_ops.knn(xd, k))
print(json.dumps({k: round(v, 2) for k, v in out.items()}))
