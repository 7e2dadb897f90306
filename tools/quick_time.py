"""Graph-replay timings of the headline ops (B=32 x 2048): chamfer fwd / fwd+bwd, EMD fwd+bwd, kNN graphs.  One JSON line."""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import losses, neighbour_ops, synthetic
from pointcloudcounterfactual_b200.structural_losses import match_cost
from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import NNDistance

dev = torch.device("cuda", 0)
a, c = (t.to(dev) for t in synthetic.s1_near(32, 2048))
rr = a.detach().requires_grad_(True)


def graph_time(fn, reps=30):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps * 1e3, 1)


def emd_fb():
    torch.autograd.grad(match_cost(rr, c).sum(), rr)


def ch_fb():
    torch.autograd.grad(losses.pykeops_chamfer(rr, c).sum(), rr)


xf = synthetic.knn_features(32, 64, 1024).to(dev)
x3 = synthetic.knn_xyz(32, 1024).to(dev)
want = sys.argv[1:] or ["chamfer", "emd", "knn"]
out = {}
if "chamfer" in want:
    out["chamfer_fwd_us"] = graph_time(lambda: NNDistance(a, c))
    out["chamfer_fwd_bwd_us"] = graph_time(ch_fb)
if "emd" in want:
    out["emd_fwd_bwd_us"] = graph_time(emd_fb, 10)
if "knn" in want:
    out["knn_feat64_k20_n1024_us"] = graph_time(lambda: neighbour_ops.knn(xf, 20))
    out["knn_xyz_k20_n1024_us"] = graph_time(lambda: neighbour_ops.knn(x3, 20))
    x25 = synthetic.knn_xyz(32, 2048).to(dev)
    out["knn_xyz_k4_n2048_us"] = graph_time(lambda: neighbour_ops.knn(x25, 4))
    out["knn_xyz_k8_n2048_us"] = graph_time(lambda: neighbour_ops.knn(x25, 8))
    out["knn_xyz_k4_n1024_us"] = graph_time(lambda: neighbour_ops.knn(x3, 4))
print(json.dumps(out))
if "gfilt" in want:
    xg = synthetic.knn_xyz(32, 2048).to(dev).requires_grad_(True)

    def gf():
        o = neighbour_ops.graph_filtering(xg, 4)
        torch.autograd.grad(o, xg, o)

    print(json.dumps({"graph_filtering_fwd_bwd_us": graph_time(gf)}))
if "knn2048" in want:
    o = {}
    for c in (64, 128):
        xx = synthetic.knn_features(32, c, 2048).to(dev)
        o[f"knn_feat{c}_k25_n2048_us"] = graph_time(lambda: neighbour_ops.knn(xx, 25), 10)
    print(json.dumps(o))
