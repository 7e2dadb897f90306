"""Graph-replay timing of get_graph_features forward and forward+backward (B=32, C=64): the bench's sub-metric."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloudcounterfactual_b200 import neighbour_ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from knn_time import ev

dev = torch.device("cuda:0")
for (B, C, N, K) in ((32, 64, 2048, 25), (32, 64, 1024, 20)):
    x = torch.randn(B, C, N, device=dev)
    idx = neighbour_ops.knn(x, K)
    xa = x.detach().requires_grad_(True)
    def f():
        neighbour_ops.get_graph_features(x, idx, K)
    def fb():
        feat = neighbour_ops.get_graph_features(xa, idx, K)[1]
        torch.autograd.grad(feat, xa, feat)
    tf, tfb = ev(f, reps=20), ev(fb, reps=20)
    gb = 2 * C * B * N * K * 4 / 1e9
    print(f"B={B} C={C} N={N} k={K}: fwd {tf:.1f} us, fwd+bwd {tfb:.1f} us, bwd = difference {tfb - tf:.1f} us ({gb / ((tfb - tf) * 1e-6):.0f} GB/s)", flush=True)
