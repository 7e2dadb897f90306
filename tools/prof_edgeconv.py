"""One fused EdgeConv layer forward+backward (for ncu): args C Cout [N] [K] [B]."""
import sys
from pathlib import Path
import torch
from torch import nn
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import edgeconv, neighbour_ops, synthetic

c, cout = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
k = int(sys.argv[4]) if len(sys.argv) > 4 else 25
b = int(sys.argv[5]) if len(sys.argv) > 5 else 32
dev = torch.device("cuda", 0)
x0 = (synthetic.knn_xyz(b, n) if c == 3 else synthetic.knn_features(b, c, n)).to(dev)
idx = neighbour_ops.knn(x0, k)
w = (torch.randn(cout, 2 * c, device=dev) / (2 * c) ** 0.5).requires_grad_(True)
gamma, beta = torch.ones(cout, device=dev).requires_grad_(True), torch.zeros(cout, device=dev).requires_grad_(True)
gout = torch.randn(b, cout, n, device=dev)
for _ in range(3):
    x = x0.detach().requires_grad_(True)
    out = edgeconv.edge_conv_max(x, idx, w, gamma, beta, None, None, edgeconv.BN_TRAIN, 0.1, 1e-5, 0.2)
    out.backward(gout)
torch.cuda.synchronize()
print("ok", float(out.sum()))
