import sys, torch
sys.path.insert(0, "/root/repo")
from pointcloudcounterfactual_b200 import neighbour_ops, synthetic
dev = torch.device("cuda", 0)
x = synthetic.knn_xyz(32, 2048).to(dev)
for _ in range(3):
    neighbour_ops.knn(x, 4)
    neighbour_ops.knn(x, 25)
torch.cuda.synchronize()
