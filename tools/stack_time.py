"""Graph-replay timing of the four chained EdgeConv layers forward+backward (bench's dgcnn_edgeconv_stack) and of one
64->64 layer; PCC_NO_SIDE_STREAM=1 keeps the backward's edge sort on the caller's stream."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloudcounterfactual_b200 import edgeconv, synthetic
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from knn_time import ev

dev = torch.device("cuda:0")
B, N, K = 32, 2048, 25

class EC(torch.nn.Module):
    def __init__(self, cin, cout, act):
        super().__init__()
        self.dense = torch.nn.Conv2d(cin, cout, kernel_size=1, bias=False)
        self.bn = torch.nn.BatchNorm2d(cout)
        self.act, self.residual = act, False

torch.manual_seed(7)
h = (64, 64, 128, 256)
enc = torch.nn.ModuleList([EC(6, h[0], None)] + [EC(2 * i, o, torch.nn.LeakyReLU(0.2, inplace=True)) for i, o in zip(h[:-1], h[1:])]).to(dev)
params = list(enc.parameters())
x = synthetic.knn_xyz(B, N).to(dev).requires_grad_(True)
def step():
    xs, t = [], x
    for layer in enc:
        t = edgeconv.fused_edge_conv(layer, t, torch.empty(0), K)[1]
        xs.append(t)
    feat = torch.cat(xs, 1)
    torch.autograd.grad(feat, [x] + params, feat)
print(f"DGCNN edge-conv stack fwd+bwd: {ev(step, reps=10):.1f} us")
