"""One-off randomized parity sweep on the GPU (not part of the test suite): feature / xyz kNN against the exact oracle on
random shapes and input styles, fused EdgeConv against the float64 oracle.  Prints one line per failure and a summary."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import oracle
from oracle import edgeconv_ref
from pointcloudcounterfactual_b200 import edgeconv, neighbour_ops, synthetic

dev = torch.device("cuda", 0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
fails = 0
cases = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    c = int(rng.choice([3, 3, 32, 64, 96, 128, 17, 160]))
    n = int(rng.integers(40, 2300))
    k = int(rng.integers(1, min(33, n)))
    b = int(rng.integers(1, 4))
    style = int(rng.integers(0, 4))
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    if c == 3:
        x = torch.randn(b, 3, n, generator=g) * torch.tensor([1.0, 0.6, 0.3]).view(1, 3, 1)
        if style == 1:
            x = (x * 16).round() / 16          # many exact ties
    else:
        x = torch.randn(b, c, n, generator=g)
        if style == 1:
            x = torch.nn.functional.leaky_relu(x, 0.2) + 3.0        # far from the origin
        elif style == 2:
            z = torch.randn(b, 4, n, generator=g)
            x = torch.randn(1, c, 4, generator=g) @ z               # low-rank manifold
            x[:, :, ::53] *= 25.0                                   # outliers
        elif style == 3:
            x = (x * 4).round() / 4                                 # exact ties in feature space
    x = x.contiguous()
    idx, dist = neighbour_ops.knn_indices(x.to(dev), k, return_dist=True)
    eidx, edist = oracle.knn(x.numpy(), k, return_dist=True)
    cases += 1
    if not (np.array_equal(idx.cpu().numpy(), eidx) and np.array_equal(dist.cpu().numpy(), edist)):
        fails += 1
        print("KNN MISMATCH", dict(c=c, n=n, k=k, b=b, style=style), flush=True)
for it in range(12):
    c = int(rng.choice([3, 8, 16, 64]))
    cout = int(rng.choice([4, 8, 12, 32, 64, 128]))
    n = int(rng.integers(30, 900))
    k = int(rng.integers(1, min(26, n)))
    b = int(rng.integers(1, 3))
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    x0 = torch.randn(b, c, n, generator=g)
    idx = torch.randint(0, n, (b, n, k), generator=g)
    w0 = torch.randn(cout, 2 * c, generator=g) / (2 * c) ** 0.5
    g0, b0 = torch.randn(cout, generator=g), torch.randn(cout, generator=g) * 0.3
    gout = torch.randn(b, cout, n, generator=g)
    slope = [None, 0.0, 0.2][int(rng.integers(0, 3))]
    xd = x0.to(dev).requires_grad_(True)
    w, gm, bt = (t.to(dev).requires_grad_(True) for t in (w0, g0, b0))
    out = edgeconv.edge_conv_max(xd, idx.to(dev), w, gm, bt, None, None, edgeconv.BN_TRAIN, 0.1, 1e-5, slope)
    out.backward(gout.to(dev))
    xr = x0.double().requires_grad_(True)
    wr, gr, br = (t.double().requires_grad_(True) for t in (w0, g0, b0))
    ref = edgeconv_ref.edge_conv_max(xr, idx, wr, gr, br, None, None, True, 0.1, 1e-5, slope)[0]
    ref.backward(gout.double())
    def rel(a, bb):
        a, bb = a.detach().cpu().double(), bb.detach().double()
        return float((a - bb).abs().max() / max(float(bb.abs().max()), 1e-30))
    errs = [rel(out, ref), rel(xd.grad, xr.grad), rel(w.grad, wr.grad), rel(gm.grad, gr.grad), rel(bt.grad, br.grad)]
    cases += 1
    if errs[0] > 2e-5 or max(errs[1:]) > 2e-4:
        fails += 1
        # diagnostic: the smallest gap between the two best DISTINCT neighbours of any (point, channel) in float64 -- a gap
        # below fp32 resolution means the arg-max itself is ambiguous in fp32 (the gradient then goes to another neighbour)
        with torch.no_grad():
            xx = x0.double()
            ws = w0.double()
            u = torch.einsum("oc,bcn->bno", ws[:, :c], xx)
            v = torch.einsum("oc,bcn->bno", ws[:, c:] - ws[:, :c], xx)
            y = torch.gather(u, 1, idx.reshape(b, n * k, 1).expand(-1, -1, cout)).view(b, n, k, cout) + v.unsqueeze(2)
            sgn = torch.where(g0 >= 0, 1.0, -1.0).double()
            ys = (y * sgn).sort(dim=2, descending=True)[0]
            gaps = (ys[:, :, :1] - ys)                       # gap to the best, per slot
            gaps = torch.where(gaps > 0, gaps, torch.full_like(gaps, 1e9)).min(dim=2)[0]
            print("   smallest positive top-2 gap:", float(gaps.min()), "scale", float(y.abs().max()))
        print("EDGECONV MISMATCH", dict(c=c, cout=cout, n=n, k=k, b=b, slope=slope), errs, flush=True)
from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import NNDistance, NNDistanceGrad  # noqa: E402

for it in range(20):
    b = int(rng.integers(1, 5))
    n, m = int(rng.integers(1, 3000)), int(rng.integers(1, 3000))
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    a, c2 = torch.randn(b, n, 3, generator=g), torch.randn(b, m, 3, generator=g)
    if it % 3 == 1:
        a, c2 = (a * 8).round() / 8, (c2 * 8).round() / 8      # exact ties: lowest index must win
    d1, i1, d2, i2 = NNDistance(a.to(dev), c2.to(dev))
    e1, j1, e2, j2 = oracle.nn_distance(a.numpy(), c2.numpy())
    g1, g2 = torch.randn(b, n, generator=g), torch.randn(b, m, generator=g)
    ga, gc = NNDistanceGrad(a.to(dev), c2.to(dev), i1, i2, g1.to(dev), g2.to(dev))
    ea, ec = oracle.nn_distance_grad(a.numpy(), c2.numpy(), j1, j2, g1.numpy(), g2.numpy())
    cases += 1
    ok = (np.array_equal(d1.cpu().numpy(), e1) and np.array_equal(i1.cpu().numpy(), j1) and np.array_equal(d2.cpu().numpy(), e2)
          and np.array_equal(i2.cpu().numpy(), j2))
    scale = max(np.abs(ea).max(), np.abs(ec).max(), 1e-30)
    ok = ok and np.abs(ga.cpu().numpy() - ea).max() / scale < 1e-5 and np.abs(gc.cpu().numpy() - ec).max() / scale < 1e-5
    if not ok:
        fails += 1
        print("CHAMFER MISMATCH", dict(b=b, n=n, m=m, it=it), flush=True)
print(f"fuzz: {cases} cases, {fails} failures")

# ---- round-2 paths: the tensor-core xyz filter (forced), the sorted gather backward, the tcgen05 GEMM ----
import os  # noqa: E402
from pointcloudcounterfactual_b200 import _lib as L  # noqa: E402

os.environ["PCC_KNN3_TC"] = "1"
for it in range(24):
    n = int(rng.integers(256, 2049))
    npad = (n + 127) // 128 * 128
    k = int(rng.integers(1, min(32, npad // 32) + 1))
    b = int(rng.integers(1, 4))
    style = int(rng.integers(0, 4))
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    x = torch.randn(b, 3, n, generator=g) * torch.tensor([1.0, 0.6, 0.3]).view(1, 3, 1)
    if style == 1:
        x = (x * 16).round() / 16                      # many exact ties
    elif style == 2:
        x = x * 1e-3 + torch.tensor([50.0, -20.0, 7.0]).view(1, 3, 1)   # tiny cloud far from the origin
    elif style == 3:
        x[:, :, ::97] *= 200.0                          # outliers stretch the fp16 scale
    pm = bool(rng.integers(0, 2))
    r0 = L.route_counts()
    if pm:
        from pointcloudcounterfactual_b200.keops import LazyTensor
        xt = x.transpose(2, 1).contiguous().to(dev)
        idx = ((LazyTensor(xt[:, :, None, :]) - LazyTensor(xt[:, None, :, :])) ** 2).sum(-1).argKmin(k, dim=2)
    else:
        idx = neighbour_ops.knn(x.to(dev), k)
    r1 = L.route_counts()
    eidx = oracle.knn(x.numpy(), k)
    cases += 1
    if r1.get("knn3_tc", 0) == r0.get("knn3_tc", 0) or not np.array_equal(idx.cpu().numpy(), eidx):
        fails += 1
        print("KNN3_TC MISMATCH / not routed", dict(n=n, k=k, b=b, style=style, pm=pm), flush=True)
del os.environ["PCC_KNN3_TC"]

# many clouds: the grids of the xyz kernels (query splits per cloud, resident-CTA fill) and of the tcgen05 kernels depend on b
for it in range(8):
    b = int(rng.choice([17, 33, 40, 64, 100]))
    c = int(rng.choice([3, 3, 64, 32]))
    n = int(rng.integers(260, 2049)) if c == 3 else int(rng.integers(64, 1200))
    k = int(rng.integers(1, 33))
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    x = torch.randn(b, c, n, generator=g)
    if c == 3 and it % 2:
        os.environ["PCC_KNN3_TC"] = "1"
    try:
        idx, dist = neighbour_ops.knn_indices(x.to(dev), k, return_dist=True)
    except RuntimeError as exc:  # forced filter outside its shapes
        idx = None
        if "unsupported" not in str(exc):
            raise
    os.environ.pop("PCC_KNN3_TC", None)
    if idx is None:
        idx, dist = neighbour_ops.knn_indices(x.to(dev), k, return_dist=True)
    eidx, edist = oracle.knn(x.numpy(), k, return_dist=True)
    cases += 1
    if not (np.array_equal(idx.cpu().numpy(), eidx) and np.array_equal(dist.cpu().numpy(), edist)):
        fails += 1
        print("MANY-CLOUD KNN MISMATCH", dict(b=b, c=c, n=n, k=k, forced_tc=bool(c == 3 and it % 2)), flush=True)

# point-major feature kNN through the KeOps expression of the reference's pykeops_knn (knn_tc2 / knn_tc / SIMT with pm = true)
from pointcloudcounterfactual_b200.keops import LazyTensor as _LT  # noqa: E402
for it in range(16):
    c = int(rng.choice([32, 64, 96, 128, 17, 160, 5]))
    n = int(rng.integers(40, 2300))
    k = int(rng.integers(1, min(33, n)))
    b = int(rng.integers(1, 4))
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    x = torch.randn(b, c, n, generator=g)
    if it % 3 == 1:
        x = torch.nn.functional.leaky_relu(x, 0.2) * 3.0 + 2.0
    elif it % 3 == 2:
        x = (x * 4).round() / 4
    xt = x.transpose(2, 1).contiguous().to(dev)
    r0 = L.route_counts()
    idx = ((_LT(xt[:, :, None, :]) - _LT(xt[:, None, :, :])) ** 2).sum(-1).argKmin(k, dim=2)
    r1 = L.route_counts()
    eidx = oracle.knn(x.contiguous().numpy(), k)
    cases += 1
    if r1.get("pm_self", 0) == r0.get("pm_self", 0) or not np.array_equal(idx.cpu().numpy(), eidx):
        fails += 1
        print("PM FEATURE KNN MISMATCH / not routed", dict(c=c, n=n, k=k, b=b), flush=True)

os.environ["PCC_KNN_BF"] = "1"   # the experimental bf16-split feature kNN (indices only)
for it in range(24):
    c = int(rng.choice([32, 64]))
    n = int(rng.integers(256, 2049))
    npad = (n + 127) // 128 * 128
    ng = npad // 16 if npad // 64 <= 16 else npad // 32
    k = int(rng.integers(1, min(32, ng) + 1))
    b = int(rng.integers(1, 4))
    style = int(rng.integers(0, 5))
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    x = torch.randn(b, c, n, generator=g)
    if style == 1:
        x = torch.nn.functional.leaky_relu(x, 0.2) + 3.0        # far from the origin
    elif style == 2:
        z = torch.randn(b, 4, n, generator=g)
        x = torch.randn(1, c, 4, generator=g) @ z               # low-rank manifold
        x[:, :, ::53] *= 25.0                                   # outliers
    elif style == 3:
        x = (x * 4).round() / 4                                 # exact ties in feature space
    elif style == 4:
        x = x * float(10.0 ** rng.integers(-6, 5))              # extreme scales
    x = x.contiguous()
    r0 = L.route_counts()
    idx = neighbour_ops.knn(x.to(dev), k)
    r1 = L.route_counts()
    eidx = oracle.knn(x.numpy(), k)
    cases += 1
    if r1.get("knn_bf", 0) == r0.get("knn_bf", 0) or not np.array_equal(idx.cpu().numpy(), eidx):
        fails += 1
        print("KNN_BF MISMATCH / not routed", dict(c=c, n=n, k=k, b=b, style=style), flush=True)
del os.environ["PCC_KNN_BF"]

for it in range(16):
    c = int(rng.choice([1, 3, 8, 20]))
    n = int(rng.integers(8, 2200))
    k = int(rng.integers(1, 33))
    b = int(rng.integers(1, 3))
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    x = torch.randn(b, c, n, generator=g).to(dev)
    idx = torch.randint(0, n, (b, n, k), generator=g)
    if it % 4 == 1:
        idx[:, :, : max(1, k // 2)] = torch.randint(0, 3, (b, n, max(1, k // 2)), generator=g)  # hubs
    idx = idx.to(dev)
    for mode, fn in ((0, neighbour_ops.get_neighbours), (1, neighbour_ops.get_graph_features)):
        w = torch.randn(b, (2 if mode else 1) * c, n, k, generator=g).to(dev)
        a = x.clone().requires_grad_(True)
        (fn(a, idx, k)[1] * w).sum().backward()
        w64 = w.double()
        ref = torch.zeros(b, c, n, dtype=torch.float64, device=dev)
        ref.scatter_add_(2, idx.view(b, 1, n * k).expand(-1, c, -1), w64[:, :c].reshape(b, c, n * k))
        if mode:
            ref += (w64[:, c:] - w64[:, :c]).sum(3)
        err = float((a.grad.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
        cases += 1
        if err > 1e-5:
            fails += 1
            print("GATHER GRAD MISMATCH", dict(c=c, n=n, k=k, b=b, mode=mode), err, flush=True)

for it in range(16):
    bt = int(rng.integers(1, 5))
    m, n, k = int(rng.integers(1, 700)), int(rng.integers(1, 300)), int(rng.integers(1, 600))
    ksplit = int(rng.choice([1, 1, 2, 4]))
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    ta, tb = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    A = torch.randn(bt, k, m, generator=g).to(dev) if ta else torch.randn(bt, m, k, generator=g).to(dev)
    Bm = torch.randn(bt, k, n, generator=g).to(dev) if tb else torch.randn(bt, n, k, generator=g).to(dev)
    sa = (A.stride(0), 1, m) if ta else (A.stride(0), k, 1)
    sb = (Bm.stride(0), 1, n) if tb else (Bm.stride(0), k, 1)
    out = torch.empty(bt * ksplit, m, n, device=dev)
    edgeconv.gemm_nt(A, sa, Bm, sb, out, (m * n, n, 1), bt, m, n, k, ksplit=ksplit)
    got = out.view(bt, ksplit, m, n).sum(1)
    if got is not None:
        A2 = A.transpose(1, 2) if ta else A
        B2 = Bm.transpose(1, 2) if tb else Bm
        ref = torch.bmm(A2.double(), B2.double().transpose(1, 2))
        err = float((got.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
        cases += 1
        if err > 3e-6:
            fails += 1
            print("GEMM MISMATCH", dict(bt=bt, m=m, n=n, k=k, ta=ta, tb=tb), err, flush=True)
print(f"fuzz (round-2 paths included): {cases} cases, {fails} failures")
