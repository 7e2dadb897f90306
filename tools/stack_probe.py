"""Where the DGCNN edge-convolution stack spends its time on chained (real) features: per layer kNN and fused layer."""
import json, sys
from pathlib import Path
import torch
from torch import nn
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import edgeconv, neighbour_ops, synthetic

dev = torch.device("cuda", 0)
B, N, K = 32, 2048, 25


class Layer(nn.Module):
    def __init__(s, cin, cout, act):
        super().__init__()
        s.dense, s.bn, s.act, s.residual = nn.Conv2d(cin, cout, 1, bias=False), nn.BatchNorm2d(cout), act, False


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 3)


torch.manual_seed(7)
h_dim = (64, 64, 128, 256)
enc = nn.ModuleList([Layer(6, 64, None)] + [Layer(2 * i, o, nn.LeakyReLU(0.2)) for i, o in zip(h_dim[:-1], h_dim[1:])]).to(dev)
h = synthetic.knn_xyz(B, N).to(dev)
res = {}
for li, layer in enumerate(enc):
    idx = neighbour_ops.knn(h, K)
    r = {"C": h.shape[1], "knn_ms": timed(lambda: neighbour_ops.knn(h, K)),
         "norm2_mean": float((h * h).sum(1).mean()), "chan_mean_abs": float(h.mean(2).abs().mean()),
         "chan_std": float(h.std(2).mean())}
    d = neighbour_ops.knn_indices(h, K, return_dist=True)[1]
    r["dk_mean"] = float(d[..., -1].mean())
    deg = torch.zeros(B, N, device=dev).scatter_add_(1, idx.reshape(B, -1), torch.ones(B, N * K, device=dev))
    r["max_indegree"] = int(deg.max())
    hh = h.detach().requires_grad_(True)

    def fb():
        out = edgeconv.fused_edge_conv(layer, hh, idx, K)[1]
        torch.autograd.grad(out, [hh] + list(layer.parameters()), out)

    r["layer_fwd_bwd_ms"] = timed(fb)
    with torch.no_grad():
        h = edgeconv.fused_edge_conv(layer, h, idx, K)[1]
    res[f"layer{li}"] = r
print(json.dumps(res))

# ---- why is the C=128 feature kNN slow on chained features?  (h = layer-3 output here; rebuild the layer-3 INPUT) ----
h = synthetic.knn_xyz(B, N).to(dev)
with torch.no_grad():
    for layer in list(enc)[:3]:
        h = edgeconv.fused_edge_conv(layer, h, torch.empty(0), K)[1]
tests = {
    "iid_c128": synthetic.knn_features(B, 128, N).to(dev),
    "real_c128": h,
    "real_centered": (h - h.mean(2, keepdim=True)).contiguous(),
    "real_first64ch": h[:, :64].contiguous(),
    "real_noise": (h + 1e-3 * torch.randn_like(h)).contiguous(),
    "real_k8": h,
}
out = {}
for name, t in tests.items():
    kk = 8 if name.endswith("k8") else K
    out[name] = {"knn_ms": timed(lambda: neighbour_ops.knn(t, kk)), "norm2_mean": float((t * t).sum(1).mean()),
                 "norm2_max": float((t * t).sum(1).max())}
    d = neighbour_ops.knn_indices(t, kk, return_dist=True)[1]
    out[name]["dk_mean"] = float(d[..., -1].mean())
    out[name]["d1_min"] = float(d[..., 1].min())
    out[name]["exact_dups"] = int((d[..., 1] == 0).sum())
print(json.dumps(out))
