"""Small, fixed launch sequence for ncu: one call of each hot operator at the BASELINE sizes."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import neighbour_ops, synthetic  # noqa: E402
from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import (  # noqa: E402
    MatchCostFused, NNDistance, NNDistanceGrad, NNDistanceTC)

dev = torch.device("cuda", 0)
B, N = 32, 2048
recon, ref = (t.to(dev) for t in synthetic.s1_near(B, N))
x3 = synthetic.knn_xyz(B, 1024).to(dev)
xf = synthetic.knn_features(B, 64, 1024).to(dev)
x25 = synthetic.knn_xyz(B, 2048).to(dev)
for _ in range(2):
    d1, i1, d2, i2 = NNDistance(recon, ref)
    g = torch.ones_like(d1) / N
    NNDistanceGrad(recon, ref, i1, i2, g, g)
    MatchCostFused(recon, ref, True, False)
    neighbour_ops.knn(x3, 20)
    neighbour_ops.knn(xf, 20)
    neighbour_ops.knn(x25, 25)   # knn3_tc_kernel (fp16 tcgen05 candidate filter)
    NNDistanceTC(recon, ref)      # nn_tc_kernel (opt-in tensor-core Chamfer search)
torch.cuda.synchronize()
print("ok")
