"""Target of the kNN ncu captures: xyz (N=1024 k=20, N=2048 k=25) and 64-dim features, two calls each."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import neighbour_ops, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
xf = synthetic.knn_features(32, 64, 1024).to(dev)
x3 = synthetic.knn_xyz(32, 1024).to(dev)
x25 = synthetic.knn_xyz(32, 2048).to(dev)
for _ in range(2):
    if which in ("all", "feat"):
        neighbour_ops.knn(xf, 20)
    if which in ("all", "xyz"):
        neighbour_ops.knn(x3, 20)
        neighbour_ops.knn(x25, 25)
torch.cuda.synchronize()
print("ok")
