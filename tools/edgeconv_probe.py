"""Times the fused EdgeConv layer against the reference op sequence (torch) on the four DGCNN layers
(C = 3, 64, 64, 128 -> Cout = 64, 64, 128, 256; N = 2048, k = 25, B = 32).  Prints one JSON line."""
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

    def forward(s, x):
        y = s.bn(s.dense(x))
        return s.act(y) if s.act is not None else y


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {}
for (c, cout) in [(3, 64), (64, 64), (64, 128), (128, 256)]:
    x0 = (synthetic.knn_xyz(B, N) if c == 3 else synthetic.knn_features(B, c, N)).to(dev)
    idx = neighbour_ops.knn(x0, K)
    layer = Layer(2 * c, cout, None if c == 3 else nn.LeakyReLU(0.2, inplace=True)).to(dev)
    gout = torch.randn(B, cout, N, device=dev)

    def fused(bwd):
        x = x0.detach().requires_grad_(True)
        out = edgeconv.fused_edge_conv(layer, x, idx, K)[1]
        if bwd:
            out.backward(gout)

    def plain(bwd):
        x = x0.detach().requires_grad_(True)
        out = layer(neighbour_ops.get_graph_features(x, idx, K)[1]).max(dim=3)[0]
        if bwd:
            out.backward(gout)

    r = {"fused_fwd_ms": timed(lambda: fused(False)), "fused_fwd_bwd_ms": timed(lambda: fused(True)),
         "torch_fwd_ms": timed(lambda: plain(False)), "torch_fwd_bwd_ms": timed(lambda: plain(True))}
    res[f"c{c}_cout{cout}"] = {k: round(v, 3) for k, v in r.items()}
print(json.dumps({"edgeconv_b32_n2048_k25": res, "cudnn_tf32_for_torch_conv": torch.backends.cudnn.allow_tf32}))
