"""Randomized GPU parity sweep, part 4: the product against the REFERENCE's own CUDA kernels (oracle/_ref, compiled unmodified)
on random shapes -- what tests/test_ref_cuda_parity.py does at fixed sizes.
  * NNDistance (bit-exact distances and indices) and NNDistanceGrad on ragged / tied / duplicated clouds
  * ApproxMatch + MatchCost + MatchCostGrad and the fused path on ragged clouds (match to 1e-6, cost / gradients to 1e-5)
Usage: python tools/fuzz_parity4.py [seed] [iterations]"""
import sys
import warnings
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import build_ref  # noqa: E402
from pointcloudcounterfactual_b200 import synthetic  # noqa: E402
from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import (  # noqa: E402
    ApproxMatch, MatchCost, MatchCostFused, MatchCostGrad, NNDistance, NNDistanceGrad)

warnings.filterwarnings("ignore")
dev = torch.device("cuda", 0)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rng = np.random.default_rng(seed)
if not build_ref.available("structural_losses_backend_ref"):
    print("oracle/_ref is not built: nothing to compare with")
    sys.exit(0)
ref = build_ref.load_ref("structural_losses_backend_ref")
fails = cases = 0


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


for it in range(iters):
    b, n, m = int(rng.integers(1, 9)), int(rng.integers(1, 4000)), int(rng.integers(1, 4000))
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    a, c = torch.randn(b, n, 3, generator=g), torch.randn(b, m, 3, generator=g) * float(rng.choice([1.0, 0.2, 4.0]))
    style = it % 4
    if style == 1:
        a, c = (a * 8).round() / 8, (c * 8).round() / 8
    elif style == 2:
        a, c = a * 1e-3 + 2.0, c * 1e-3 + 2.0
    elif style == 3 and m > 8:
        c[:, : m // 2] = c[:, m // 2: m // 2 + m // 2]
    ta, tc = a.to(dev), c.to(dev)
    rd1, ri1, rd2, ri2 = ref.NNDistance(ta, tc)
    d1, i1, d2, i2 = NNDistance(ta, tc)
    gd1, gd2 = torch.randn(b, n, generator=g).to(dev), torch.randn(b, m, generator=g).to(dev)
    torch.cuda.synchronize()  # the reference memsets on the legacy stream (nndistance.cu:150-151)
    rg1, rg2 = ref.NNDistanceGrad(ta, tc, ri1, ri2, gd1, gd2)
    torch.cuda.synchronize()
    g1, g2 = NNDistanceGrad(ta, tc, i1, i2, gd1, gd2)
    cases += 1
    ok = torch.equal(i1, ri1) and torch.equal(i2, ri2) and torch.equal(d1, rd1) and torch.equal(d2, rd2)
    e = (rel(g1, rg1), rel(g2, rg2))
    if not ok or max(e) > 1e-5:
        fails += 1
        print("NN_DISTANCE vs reference CUDA MISMATCH", dict(b=b, n=n, m=m, style=style), ok, e, flush=True)

for it in range(iters):
    b, n, m = int(rng.integers(1, 5)), int(rng.integers(2, 1800)), int(rng.integers(2, 1800))
    a, c = synthetic.s2_far(b, n, m) if it % 2 else synthetic.s1_near(b, n)
    ta, tc = a.to(dev), c.to(dev)
    rmatch, _ = ref.ApproxMatch(ta, tc)
    rcost = ref.MatchCost(ta, tc, rmatch)
    rg1, rg2 = ref.MatchCostGrad(ta, tc, rmatch)
    torch.cuda.synchronize()
    match, _ = ApproxMatch(ta, tc)
    cost = MatchCost(ta, tc, match)
    g1, g2 = MatchCostGrad(ta, tc, match)
    fc, f1, f2 = MatchCostFused(ta, tc)
    cases += 1
    e = (float((match - rmatch).abs().max()) / max(1.0, float(rmatch.max())), rel(cost, rcost), rel(g1, rg1), rel(g2, rg2),
         rel(fc, rcost), rel(f1, rg1), rel(f2, rg2))
    if e[0] > 1e-6 or max(e[1:]) > 1e-5:
        fails += 1
        print("APPROXMATCH vs reference CUDA MISMATCH", dict(b=b, n=int(ta.shape[1]), m=int(tc.shape[1])), e, flush=True)
# many clouds (the generate loop runs B = 256): small clouds, large batch
for it in range(max(2, iters // 3)):
    b, n, m = int(rng.choice([33, 64, 100, 256])), int(rng.integers(2, 300)), int(rng.integers(2, 300))
    a, c = synthetic.s2_far(b, n, m)
    ta, tc = a.to(dev), c.to(dev)
    rd1, ri1, rd2, ri2 = ref.NNDistance(ta, tc)
    d1, i1, d2, i2 = NNDistance(ta, tc)
    rmatch, _ = ref.ApproxMatch(ta, tc)
    rcost = ref.MatchCost(ta, tc, rmatch)
    rg1, rg2 = ref.MatchCostGrad(ta, tc, rmatch)
    torch.cuda.synchronize()
    fc, f1, f2 = MatchCostFused(ta, tc)
    cases += 1
    ok = torch.equal(i1, ri1) and torch.equal(i2, ri2) and torch.equal(d1, rd1) and torch.equal(d2, rd2)
    e = (rel(fc, rcost), rel(f1, rg1), rel(f2, rg2))
    if not ok or max(e) > 1e-5:
        fails += 1
        print("MANY-CLOUD vs reference CUDA MISMATCH", dict(b=b, n=n, m=m), ok, e, flush=True)
print(f"fuzz4 seed {seed}: {cases} cases, {fails} failures")
