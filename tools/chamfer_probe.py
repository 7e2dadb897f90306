"""Chamfer probe (GPU box): forward and forward+backward timings under CUDA-graph replay, B=32 x 2048."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import losses, synthetic  # noqa: E402
from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import NNDistance, NNDistanceGrad  # noqa: E402

dev = torch.device("cuda", 0)
B, N = 32, 2048


def ev(fn, reps=50, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def graph(fn, reps=50):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return ev(g.replay, reps)


recon, ref = (t.to(dev) for t in synthetic.s1_near(B, N))
out = {"nndistance_us": round(graph(lambda: NNDistance(recon, ref)), 1)}
d1, i1, d2, i2 = NNDistance(recon, ref)
g1 = torch.ones_like(d1) / N
out["nndistancegrad_us"] = round(graph(lambda: NNDistanceGrad(recon, ref, i1, i2, g1, g1)), 1)
rr = recon.detach().requires_grad_(True)


def fb():
    loss = losses.pykeops_chamfer(rr, ref)
    torch.autograd.grad(loss.sum(), rr)


out["pykeops_chamfer_fwd_bwd_us"] = round(graph(fb), 1)
if hasattr(losses, "chamfer_fused"):
    def fb2():
        loss = losses.chamfer_fused(rr, ref)
        torch.autograd.grad(loss.sum(), rr)
    out["chamfer_fused_fwd_bwd_us"] = round(graph(fb2), 1)
print(json.dumps(out))
