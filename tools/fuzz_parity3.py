"""Randomized GPU parity sweep, part 3 (not part of the test suite):
  * fused EdgeConv layer in all three normalisation modes, large shapes (N above the staging limit, Cout up to 256, k up to 64,
    hub graphs), vs the float64 oracle
  * NNDistanceTC (tensor-core Chamfer search) on ragged clouds vs the C oracle, bit for bit
  * generic kNN: any C, k up to 128, vs the C oracle, bit for bit
  * VQ nearest codeword (argmin_small route) vs float64 brute force
Usage: python tools/fuzz_parity3.py [seed] [iterations]"""
import sys
import warnings
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import oracle  # noqa: E402
from oracle import edgeconv_ref  # noqa: E402
from pointcloudcounterfactual_b200 import edgeconv, keops, neighbour_ops  # noqa: E402
from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import NNDistanceTC  # noqa: E402

warnings.filterwarnings("ignore")
dev = torch.device("cuda", 0)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rng = np.random.default_rng(seed)
fails = cases = 0


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def gen():
    return torch.Generator().manual_seed(int(rng.integers(1 << 30)))


for it in range(iters):
    c = int(rng.choice([3, 4, 16, 64, 128]))
    cout = int(rng.choice([4, 8, 64, 128, 256]))
    n = int(rng.integers(30, 3000))
    k = int(rng.integers(1, min(65, n)))
    b = int(rng.integers(1, 3))
    if b * n * k * cout > 1.2e8:
        k = max(1, int(1.2e8 / (b * n * cout)))
    mode = int(rng.integers(0, 3))  # 0 eval, 1 train, 2 affine
    g = gen()
    x0 = torch.randn(b, c, n, generator=g)
    idx = torch.randint(0, n, (b, n, k), generator=g)
    if it % 3 == 1:
        idx[:, :, : max(1, k // 3)] = torch.randint(0, 5, (b, n, max(1, k // 3)), generator=g)  # hubs
    w0 = torch.randn(cout, 2 * c, generator=g) / (2 * c) ** 0.5
    g0, b0 = torch.randn(cout, generator=g), torch.randn(cout, generator=g) * 0.3
    rm0, rv0 = torch.randn(cout, generator=g) * 0.2, torch.rand(cout, generator=g) + 0.5
    gout = torch.randn(b, cout, n, generator=g)
    slope = [None, 0.0, 0.2][int(rng.integers(0, 3))]
    xd = x0.to(dev).requires_grad_(True)
    w, gm, bt = (t.to(dev).requires_grad_(True) for t in (w0, g0, b0))
    rm, rv = rm0.to(dev).clone(), rv0.to(dev).clone()
    xr = x0.double().requires_grad_(True)
    wr, gr, br = (t.double().requires_grad_(True) for t in (w0, g0, b0))
    if mode == 2:
        out = edgeconv.edge_conv_max(xd, idx.to(dev), w, None, bt, None, None, edgeconv.AFFINE, 0.1, 1e-5, slope)
        # no normalisation: y = conv + bias
        ref = edgeconv_ref.edge_conv_max(xr, idx, wr, torch.ones(cout, dtype=torch.float64), br, torch.zeros(cout, dtype=torch.float64),
                                         torch.ones(cout, dtype=torch.float64) - 1e-5, False, 0.1, 1e-5, slope)[0]
        params = [(bt, br), (w, wr)]
    else:
        out = edgeconv.edge_conv_max(xd, idx.to(dev), w, gm, bt, rm, rv, mode, 0.1, 1e-5, slope)
        ref = edgeconv_ref.edge_conv_max(xr, idx, wr, gr, br, rm0.double(), rv0.double(), mode == 1, 0.1, 1e-5, slope)[0]
        params = [(gm, gr), (bt, br), (w, wr)]
    out.backward(gout.to(dev))
    ref.backward(gout.double())
    errs = [rel(out, ref), rel(xd.grad, xr.grad)] + [rel(p.grad, q.grad) for p, q in params]
    cases += 1
    if errs[0] > 3e-5 or max(errs[1:]) > 3e-4:
        # Where does the input gradient differ?  Isolated elements = decisions fp32 cannot resolve (two neighbours tied for the
        # maximum to within rounding, or a winning pre-activation within rounding of zero where the activation's slope jumps);
        # only a difference spread over MANY elements counts as a failure.
        d = (xd.grad.detach().cpu().double() - xr.grad).abs()
        scale = float(xr.grad.abs().max())
        n_off = int((d > 1e-4 * scale).sum())
        med = float(d.flatten().median()) / scale
        with torch.no_grad():
            xx, ws = x0.double(), w0.double()
            u = torch.einsum("oc,bcn->bno", ws[:, :c], xx)
            v = torch.einsum("oc,bcn->bno", ws[:, c:] - ws[:, :c], xx)
            y = torch.gather(u, 1, idx.reshape(b, n * k, 1).expand(-1, -1, cout)).view(b, n, k, cout) + v.unsqueeze(2)
            if mode == 1:
                mu, var = y.mean((0, 1, 2)), y.var((0, 1, 2), unbiased=False)
                z = (y - mu) / torch.sqrt(var + 1e-5) * g0.double() + b0.double()
            elif mode == 0:
                z = (y - rm0.double()) / torch.sqrt(rv0.double() + 1e-5) * g0.double() + b0.double()
            else:
                z = y + b0.double()
            sgn = torch.ones(cout, dtype=torch.float64) if mode == 2 else torch.where(g0 >= 0, 1.0, -1.0).double()
            ys = y * sgn
            top = ys.max(2, keepdim=True)[0]
            gaps = top - ys
            zs = float(z.abs().max())
            ties = int(((gaps > 0) & (gaps < 2e-6 * float(ys.abs().max()))).any(2).sum())
            zmax = torch.gather(z, 2, ys.argmax(2, keepdim=True)).squeeze(2)
            zeros = int((zmax.abs() < 2e-6 * zs).sum()) if slope is not None else 0
        isolated = med < 1e-6 and n_off <= 4 * c * max(1, ties + zeros) + 4 * (ties + zeros) * k and (ties + zeros) > 0
        info = dict(c=c, cout=cout, n=n, k=k, b=b, mode=mode, slope=slope, hubs=it % 3 == 1)
        if isolated:
            print("edgeconv: fp32-unresolvable decisions", info, f"median err {med:.1e}, {n_off} of {d.numel()} elements off, "
                  f"{ties} near-tied maxima, {zeros} winners at the activation's kink", flush=True)
        else:
            fails += 1
            print("EDGECONV MISMATCH", info, errs, f"median err {med:.1e}, {n_off} elements off, ties {ties}, kinks {zeros}", flush=True)

for it in range(iters):
    b, n, m = int(rng.integers(1, 5)), int(rng.integers(256, 2561)), int(rng.integers(256, 2561))
    g = gen()
    a, c = torch.randn(b, n, 3, generator=g), torch.randn(b, m, 3, generator=g) * float(rng.choice([1.0, 0.3, 3.0]))
    style = it % 4
    if style == 1:
        a, c = (a * 8).round() / 8, (c * 8).round() / 8
    elif style == 2:
        a, c = a * 1e-2 + 5.0, c * 1e-2 + 5.0   # small clouds far from the origin
    elif style == 3:
        c[:, : m // 2] = c[:, m // 2: m // 2 + m // 2]
    d1, i1, d2, i2 = NNDistanceTC(a.to(dev), c.to(dev))
    e1, j1, e2, j2 = oracle.nn_distance(a.numpy(), c.numpy())
    cases += 1
    if not (np.array_equal(d1.cpu().numpy(), e1) and np.array_equal(i1.cpu().numpy(), j1) and
            np.array_equal(d2.cpu().numpy(), e2) and np.array_equal(i2.cpu().numpy(), j2)):
        fails += 1
        print("NN_TC MISMATCH", dict(b=b, n=n, m=m, style=style), flush=True)

for it in range(iters):
    c = int(rng.choice([1, 2, 5, 7, 24, 33, 48, 200]))
    n = int(rng.integers(2, 3000))
    k = int(rng.integers(1, min(129, n + 1)))
    b = int(rng.integers(1, 3))
    g = gen()
    x = torch.randn(b, c, n, generator=g)
    if it % 3 == 1:
        x = (x * 4).round() / 4
    idx, dist = neighbour_ops.knn_indices(x.to(dev), k, return_dist=True)
    eidx, edist = oracle.knn(x.numpy(), k, return_dist=True)
    cases += 1
    if not (np.array_equal(idx.cpu().numpy(), eidx) and np.array_equal(dist.cpu().numpy(), edist)):
        fails += 1
        print("GENERIC KNN MISMATCH", dict(c=c, n=n, k=k, b=b), flush=True)

for it in range(iters):
    b, ncodes, book, dim = int(rng.integers(1, 600)), int(rng.integers(1, 9)), int(rng.integers(1, 65)), int(rng.integers(1, 17))
    g = gen()
    q, r = torch.randn(b, ncodes, dim, generator=g), torch.randn(b, book, dim, generator=g)
    idx = keops.argkmin(q.to(dev), r.to(dev), 1).cpu().numpy()[..., 0]
    d = ((q.double()[:, :, None, :] - r.double()[:, None, :, :]) ** 2).sum(-1).numpy()
    want = d.argmin(2)
    cases += 1
    if not np.array_equal(idx, want):
        dg, dw = np.take_along_axis(d, idx[..., None].astype(np.int64), 2), np.take_along_axis(d, want[..., None], 2)
        if np.abs(dg - dw).max() > 1e-5 * max(1.0, float(dw.max())):
            fails += 1
            print("VQ ARGMIN MISMATCH", dict(b=b, ncodes=ncodes, book=book, dim=dim), flush=True)
print(f"fuzz3 seed {seed}: {cases} cases, {fails} failures")
