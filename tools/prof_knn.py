import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import neighbour_ops, synthetic
dev = torch.device("cuda", 0)
xf = synthetic.knn_features(32, 64, 1024).to(dev)
x3 = synthetic.knn_xyz(32, 1024).to(dev)
for _ in range(2):
    neighbour_ops.knn(xf, 20)
    neighbour_ops.knn(x3, 20)
torch.cuda.synchronize()
print("ok")
