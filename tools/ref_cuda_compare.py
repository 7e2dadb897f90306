"""Side-by-side timing of the REFERENCE's own CUDA kernels (oracle/_ref, compiled unmodified from /root/reference) and
this library on the same B200, BASELINE config 3 shapes (B=32, N=2048).  Measurement tooling, not part of the product
path: it is the only place outside tests/ that loads oracle/_ref, and bench.py never does.

    python tools/ref_cuda_compare.py > gpurun_out/ref_cuda_compare.json
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import build_ref  # noqa: E402
from pointcloudcounterfactual_b200 import synthetic  # noqa: E402
from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import (  # noqa: E402
    ApproxMatch, MatchCost, MatchCostFused, MatchCostGrad, NNDistance, NNDistanceGrad)

dev = torch.device("cuda", 0)
B, N = 32, 2048


def ev(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps * 1e3, 1)  # microseconds


out = {"shape": {"batch": B, "points": N}, "unit": "us", "timing": "CUDA events, eager launches, mean of 10"}
if not build_ref.available("structural_losses_backend_ref"):
    print(json.dumps({"unavailable": "oracle/_ref/structural not built"}))
    sys.exit(0)
ref = build_ref.load_ref("structural_losses_backend_ref")
a, c = (t.to(dev) for t in synthetic.s1_near(B, N))
g1 = torch.full((B, N), 1.0 / N, device=dev)

rd1, ri1, rd2, ri2 = ref.NNDistance(a, c)
out["nn_distance_fwd"] = {"reference_cuda": ev(lambda: ref.NNDistance(a, c)), "b200": ev(lambda: NNDistance(a, c))}
out["nn_distance_bwd"] = {"reference_cuda": ev(lambda: ref.NNDistanceGrad(a, c, ri1, ri2, g1, g1)),
                          "b200": ev(lambda: NNDistanceGrad(a, c, ri1, ri2, g1, g1))}


def ref_emd():
    match, _ = ref.ApproxMatch(a, c)
    ref.MatchCost(a, c, match)
    ref.MatchCostGrad(a, c, match)


def our_emd_chain():
    match, _ = ApproxMatch(a, c)
    MatchCost(a, c, match)
    MatchCostGrad(a, c, match)


out["emd_approxmatch_matchcost_grad"] = {"reference_cuda": ev(ref_emd, reps=3, warm=1),
                                         "b200_same_chain": ev(our_emd_chain, reps=5, warm=1),
                                         "b200_fused_match_cost": ev(lambda: MatchCostFused(a, c, True, True), reps=5, warm=1)}
# ---- auction EMD (external/emd): B=32, N=2048, eps=0.005, 50 iterations (the reference's documented setting) ----
if build_ref.available("emd_backend_ref"):
    from pointcloudcounterfactual_b200.emd import emdModule

    refe = build_ref.load_ref("emd_backend_ref")
    ua, uc = (t.to(dev) for t in synthetic.auction_clouds(B, N))

    def buf(shape, dtype, fill=0):
        return torch.full(shape, fill, dtype=dtype, device=dev)

    def ref_auction():  # buffers initialised per call exactly as emd_module.py:34-45 does
        args = [ua, uc, buf((B, N), torch.float32), buf((B, N), torch.int32, -1), buf((B, N), torch.float32),
                buf((B, N), torch.int32, -1), buf((B, N), torch.int32), buf((B, N), torch.float32),
                buf((B, N), torch.float32), buf((B * N,), torch.int32), buf((512,), torch.int32),
                buf((512,), torch.int32), buf((512,), torch.int32), buf((B * N,), torch.int32), 0.005, 50]
        refe.forward(*args)

    mod = emdModule()
    out["auction_emd_fwd_eps0.005_iters50"] = {"reference_cuda": ev(ref_auction, reps=3, warm=1),
                                               "b200": ev(lambda: mod(ua, uc, 0.005, 50), reps=5, warm=1)}
for k, v in out.items():
    if isinstance(v, dict) and "reference_cuda" in v:
        base = v["reference_cuda"]
        v["speedup"] = {n: round(base / t, 2) for n, t in v.items() if n != "reference_cuda" and isinstance(t, float)}
print(json.dumps(out))
