"""Times get_graph_features forward+backward (B=32, C=64, N=2048, k=25) -- run under gpurun."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloudcounterfactual_b200 import neighbour_ops, _lib as L

dev = torch.device("cuda:0")
def ev_time(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3
for (B, C, N, K) in ((32, 64, 2048, 25), (32, 64, 2048, 20), (32, 64, 1024, 20), (32, 128, 2048, 25)):
    x = torch.randn(B, C, N, device=dev)
    idx = neighbour_ops.knn(x, K)  # feature-space graph: real in-degree distribution (hubs)
    g = torch.randn(B, 2 * C, N, K, device=dev)
    gx = torch.empty(B, C, N, device=dev)
    lib = L.load()
    print("max in-degree", int(torch.bincount(idx[0].flatten(), minlength=N).max()))
    def bwd():
        L.check(lib.pcc_graph_gather_grad(B, C, N, K, L.ptr(idx), 1, L.ptr(g), L.ptr(gx), L.stream_of(g)), "gg")
    us = ev_time(bwd)
    gb = g.numel() * 4 / 1e9
    print(f"B={B} C={C} N={N} k={K}: gather_grad {us:.1f} us  ({gb / (us * 1e-6):.0f} GB/s of gradient read)", flush=True)
