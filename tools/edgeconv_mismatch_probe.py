"""Replays the two EdgeConv cases the fuzz sweep (tools/fuzz_parity3.py, seed 2) flagged and shows WHERE the input-gradient
differs from the float64 oracle: isolated elements (arg-max / ReLU decisions that fp32 cannot resolve) or everywhere."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import edgeconv_ref
from pointcloudcounterfactual_b200 import edgeconv

dev = torch.device("cuda", 0)
rng = np.random.default_rng(2)
def gen():
    return torch.Generator().manual_seed(int(rng.integers(1 << 30)))
for it in range(10):
    c = int(rng.choice([3, 4, 16, 64, 128])); cout = int(rng.choice([4, 8, 64, 128, 256])); n = int(rng.integers(30, 3000))
    k = int(rng.integers(1, min(65, n))); b = int(rng.integers(1, 3))
    if b * n * k * cout > 1.2e8: k = max(1, int(1.2e8 / (b * n * cout)))
    mode = int(rng.integers(0, 3)); g = gen()
    x0 = torch.randn(b, c, n, generator=g); idx = torch.randint(0, n, (b, n, k), generator=g)
    if it % 3 == 1: idx[:, :, : max(1, k // 3)] = torch.randint(0, 5, (b, n, max(1, k // 3)), generator=g)
    w0 = torch.randn(cout, 2 * c, generator=g) / (2 * c) ** 0.5
    g0, b0 = torch.randn(cout, generator=g), torch.randn(cout, generator=g) * 0.3
    rm0, rv0 = torch.randn(cout, generator=g) * 0.2, torch.rand(cout, generator=g) + 0.5
    gout = torch.randn(b, cout, n, generator=g); slope = [None, 0.0, 0.2][int(rng.integers(0, 3))]
    if mode != 1 or k < 40: continue
    xd = x0.to(dev).requires_grad_(True); w, gm, bt = (t.to(dev).requires_grad_(True) for t in (w0, g0, b0))
    out = edgeconv.edge_conv_max(xd, idx.to(dev), w, gm, bt, rm0.to(dev).clone(), rv0.to(dev).clone(), 1, 0.1, 1e-5, slope)
    out.backward(gout.to(dev))
    xr = x0.double().requires_grad_(True); wr, gr, br = (t.double().requires_grad_(True) for t in (w0, g0, b0))
    ref = edgeconv_ref.edge_conv_max(xr, idx, wr, gr, br, rm0.double(), rv0.double(), True, 0.1, 1e-5, slope)[0]
    ref.backward(gout.double())
    d = (xd.grad.cpu().double() - xr.grad).abs(); scale = float(xr.grad.abs().max())
    flat = d.flatten().sort(descending=True)[0]
    print(dict(c=c, cout=cout, n=n, k=k, b=b, slope=slope), "max rel", float(flat[0]) / scale, "10th", float(flat[9]) / scale,
          "100th", float(flat[99]) / scale, "median", float(flat[len(flat) // 2]) / scale, "elements above 1e-4:", int((d > 1e-4 * scale).sum()), "of", d.numel())
    # decisions fp32 cannot resolve: pre-activations within 1e-6 of zero (ReLU) / top-2 gaps below 1e-6 (arg-max)
    with torch.no_grad():
        xx, ws = x0.double(), w0.double()
        u = torch.einsum("oc,bcn->bno", ws[:, :c], xx); v = torch.einsum("oc,bcn->bno", ws[:, c:] - ws[:, :c], xx)
        y = torch.gather(u, 1, idx.reshape(b, n * k, 1).expand(-1, -1, cout)).view(b, n, k, cout) + v.unsqueeze(2)
        mu, var = y.mean((0, 1, 2)), y.var((0, 1, 2), unbiased=False)
        z = (y - mu) / torch.sqrt(var + 1e-5) * g0.double() + b0.double()
        zmax = z.max(2)[0]
        print("   |pre-activation of the winner| < 1e-6:", int((zmax.abs() < 1e-6).sum()), " top-2 gaps < 1e-6 (distinct neighbours):",
              int(((zmax.unsqueeze(2) - z) > 0).logical_and((zmax.unsqueeze(2) - z) < 1e-6).any(2).sum()))
