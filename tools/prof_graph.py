"""Target of the EdgeConv front-end captures: get_graph_features forward + backward, B=32, C=64, N=2048, k=25."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import neighbour_ops, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
x = synthetic.knn_features(32, 64, 2048, seed=3000).to(dev).requires_grad_(True)
idx = neighbour_ops.knn(x.detach(), 25)
for _ in range(2):
    feat = neighbour_ops.get_graph_features(x, idx, 25)[1]
    torch.autograd.grad(feat, x, torch.ones_like(feat))
torch.cuda.synchronize()
print("ok")
