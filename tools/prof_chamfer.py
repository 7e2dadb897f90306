"""Target of the Chamfer ncu captures: NNDistance forward + backward, B=32 x 2048 (S1), two calls."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import synthetic  # noqa: E402
from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import NNDistance, NNDistanceGrad  # noqa: E402

dev = torch.device("cuda", 0)
recon, ref = (t.to(dev) for t in synthetic.s1_near(32, 2048))
for _ in range(2):
    d1, i1, d2, i2 = NNDistance(recon, ref)
    g = torch.ones_like(d1) / 2048
    NNDistanceGrad(recon, ref, i1, i2, g, g)
torch.cuda.synchronize()
print("ok")
