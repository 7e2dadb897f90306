"""One fused EMD forward+backward (B=32, N=2048, S1) -- the target of the ncu launch list."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import synthetic  # noqa: E402
from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import MatchCostFused  # noqa: E402

dev = torch.device("cuda", 0)
kind = sys.argv[1] if len(sys.argv) > 1 else "s1"
a, c = (synthetic.s1_near(32, 2048) if kind == "s1" else synthetic.s2_far(32, 2048, 2048))
a, c = a.to(dev), c.to(dev)
for _ in range(2):
    cost, g1, _ = MatchCostFused(a, c, True, False)
torch.cuda.synchronize()
print(float(cost.sum()))
