"""Randomized GPU parity sweep, part 2 (not part of the test suite): the operators fuzz_parity.py does not touch.
  * Chamfer losses (pykeops_chamfer / torch_chamfer) forward + backward on ragged, duplicated and tiny clouds vs a float64
    torch composition with the oracle's nearest-neighbour indices
  * two-operand argKmin (KeOps shim, q != r) vs float64 brute force with a near-tie guard
  * graph_filtering forward + backward vs the reference's torch composition in float64 (same kNN graph)
  * graph_max_pooling / get_local_covariance vs their torch compositions
  * match_cost (approxmatch EMD) forward + backward vs the C oracle on small ragged clouds (AMP_TOL, see tests/test_gpu_emd.py)
  * the EMD materialised vs fused paths at random sizes (fp32 bar)
  * auction EMD forward vs the C oracle
Usage: python tools/fuzz_parity2.py [seed] [iterations]"""
import sys
import warnings
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import oracle  # noqa: E402
from pointcloudcounterfactual_b200 import keops, losses, neighbour_ops, synthetic  # noqa: E402
from pointcloudcounterfactual_b200.emd import emdModule  # noqa: E402
from pointcloudcounterfactual_b200.structural_losses import match_cost  # noqa: E402
from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import (  # noqa: E402
    ApproxMatch, MatchCost, MatchCostFused, MatchCostGrad)

warnings.filterwarnings("ignore")
dev = torch.device("cuda", 0)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rng = np.random.default_rng(seed)
fails = cases = 0


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(float(np.abs(b).max()), 1e-30))


def bad(what, info, *errs):
    global fails
    fails += 1
    print(what, info, errs, flush=True)


def gen():
    return torch.Generator().manual_seed(int(rng.integers(1 << 30)))


# ---- Chamfer losses -------------------------------------------------------------------------------------------------
for it in range(iters):
    b, n, m = int(rng.integers(1, 5)), int(rng.integers(1, 2600)), int(rng.integers(1, 2600))
    g = gen()
    a, c = torch.randn(b, n, 3, generator=g), torch.randn(b, m, 3, generator=g) * 0.8 + 0.1
    style = it % 4
    if style == 1:
        a, c = (a * 8).round() / 8, (c * 8).round() / 8
    elif style == 2:
        a, c = a * 1e-3, c * 1e-3
    elif style == 3 and n > 4:
        a[:, : n // 2] = a[:, n // 2: n // 2 + n // 2]  # duplicated points
    _, i1, _, i2 = oracle.nn_distance(a.numpy(), c.numpy())
    for name, fn, mean in (("pykeops_chamfer", losses.pykeops_chamfer, True), ("torch_chamfer", losses.torch_chamfer, False)):
        ad, cd = a.to(dev).requires_grad_(True), c.to(dev).requires_grad_(True)
        w = torch.randn(b, generator=g)
        loss = fn(ad, cd)
        (loss * w.to(dev)).sum().backward()
        a64, c64 = a.double().requires_grad_(True), c.double().requires_grad_(True)
        j1 = torch.from_numpy(i1.astype(np.int64)).unsqueeze(-1).expand(-1, -1, 3)
        j2 = torch.from_numpy(i2.astype(np.int64)).unsqueeze(-1).expand(-1, -1, 3)
        d1 = ((a64 - torch.gather(c64, 1, j1)) ** 2).sum(-1)
        d2 = ((c64 - torch.gather(a64, 1, j2)) ** 2).sum(-1)
        ref = d1.mean(1) + d2.mean(1) if mean else d1.sum(1) + d2.sum(1)
        (ref * w.double()).sum().backward()
        cases += 1
        e = (rel(loss.detach().cpu(), ref.detach()), rel(ad.grad.cpu(), a64.grad), rel(cd.grad.cpu(), c64.grad))
        if max(e) > 2e-5:
            bad("CHAMFER LOSS MISMATCH", dict(fn=name, b=b, n=n, m=m, style=style), *e)

# ---- two-operand argKmin through the KeOps shim -----------------------------------------------------------------------
for it in range(iters):
    b, c = int(rng.integers(1, 4)), int(rng.choice([1, 2, 3, 4, 8, 16, 64, 100]))
    nq, nr = int(rng.integers(1, 700)), int(rng.integers(1, 900))
    k = int(rng.integers(1, min(nr, 40) + 1))
    g = gen()
    q, r = torch.randn(b, nq, c, generator=g), torch.randn(b, nr, c, generator=g)
    idx = keops.argkmin(q.to(dev), r.to(dev), k).cpu().numpy()
    d = ((q.double()[:, :, None, :] - r.double()[:, None, :, :]) ** 2).sum(-1).numpy()
    want = np.argsort(d, axis=2, kind="stable")[:, :, :k]
    cases += 1
    if not np.array_equal(idx, want):
        # float64 order vs the fp32 chain can differ only at near-ties: compare the selected distances instead
        dg = np.take_along_axis(d, idx.astype(np.int64), 2)
        dw = np.take_along_axis(d, want, 2)
        if np.abs(dg - dw).max() > 1e-5 * max(1.0, float(dw.max())):
            bad("ARGKMIN MISMATCH", dict(b=b, c=c, nq=nq, nr=nr, k=k), float(np.abs(dg - dw).max()))

# ---- graph_filtering ---------------------------------------------------------------------------------------------------
def filtering_ref(x, idx, k):  # neighbour_ops.py:122-133 written out in float64 on a given graph
    b, c, n = x.shape
    nb = torch.gather(x, 2, idx[:, :, 1:].reshape(b, 1, n * (k - 1)).expand(-1, c, -1)).view(b, c, n, k - 1)
    diff = x.unsqueeze(3) - nb
    dist = torch.sqrt((diff ** 2).sum(1))
    sigma = torch.clamp(dist[:, :, 0].mean(1, keepdim=True), min=0.005)
    w = torch.exp(-dist / sigma.unsqueeze(2))
    return (1 + w.sum(2, keepdim=False).unsqueeze(1)) * x - (w.unsqueeze(1) * nb).sum(3)


for it in range(iters):
    b, n, k = int(rng.integers(1, 5)), int(rng.integers(6, 3000)), int(rng.integers(2, 9))
    g = gen()
    x = torch.randn(b, 3, n, generator=g) * torch.tensor([1.0, 0.5, 0.25]).view(1, 3, 1)
    xd = x.to(dev).requires_grad_(True)
    out = neighbour_ops.graph_filtering(xd, k)
    gout = torch.randn(out.shape, generator=g)
    out.backward(gout.to(dev))
    idx = torch.from_numpy(oracle.knn(x.numpy(), k))
    x64 = x.double().requires_grad_(True)
    ref = filtering_ref(x64, idx, k)
    ref.backward(gout.double())
    cases += 1
    e = (rel(out.detach().cpu(), ref.detach()), rel(xd.grad.cpu(), x64.grad))
    if e[0] > 1e-5 or e[1] > 2e-4:
        bad("GRAPH FILTERING MISMATCH", dict(b=b, n=n, k=k), *e)

# ---- graph_max_pooling / get_local_covariance --------------------------------------------------------------------------
for it in range(iters):
    b, c = int(rng.integers(1, 4)), int(rng.choice([3, 4, 8, 12, 64, 6]))
    n, k = int(rng.integers(20, 1500)), int(rng.integers(1, 17))
    g = gen()
    x = torch.randn(b, c, n, generator=g)
    idx = torch.randint(0, n, (b, n, k), generator=g)
    xd, idd = x.to(dev), idx.to(dev)
    got = neighbour_ops.graph_max_pooling(xd, idd, k).cpu()
    nb = torch.gather(x, 2, idx.view(b, 1, n * k).expand(-1, c, -1)).view(b, c, n, k)
    cases += 1
    if not torch.equal(got, nb.max(3)[0]):
        bad("GRAPH MAX POOLING MISMATCH", dict(b=b, c=c, n=n, k=k), rel(got, nb.max(3)[0]))
    if c == 3:
        cov = neighbour_ops.get_local_covariance(xd, idd, k).cpu()
        nb0 = nb - nb.mean(3, keepdim=True)
        want = torch.cat([x, torch.matmul(nb0.transpose(1, 2), nb0.permute(0, 2, 3, 1)).flatten(start_dim=2).transpose(1, 2)], 1)
        cases += 1
        if rel(cov, want) > 2e-5:
            bad("LOCAL COVARIANCE MISMATCH", dict(b=b, n=n, k=k), rel(cov, want))

# ---- approxmatch EMD ---------------------------------------------------------------------------------------------------
for it in range(max(4, iters // 2)):
    b, n, m = int(rng.integers(1, 4)), int(rng.integers(2, 400)), int(rng.integers(2, 400))
    a, c = synthetic.s2_far(b, n, m)
    ematch, _ = oracle.approxmatch(a.numpy(), c.numpy())
    ecost = oracle.matchcost(a.numpy(), c.numpy(), ematch)
    eg1, eg2 = oracle.matchcostgrad(a.numpy(), c.numpy(), ematch)
    ad, cd = a.to(dev), c.to(dev)
    match, _ = ApproxMatch(ad, cd)
    cost = MatchCost(ad, cd, match)
    g1, g2 = MatchCostGrad(ad, cd, match)
    fc, f1, f2 = MatchCostFused(ad, cd)
    cases += 1
    e = (rel(cost.cpu(), ecost), rel(g1.cpu(), eg1), rel(g2.cpu(), eg2), rel(fc.cpu(), cost.cpu()), rel(f1.cpu(), g1.cpu()),
         rel(f2.cpu(), g2.cpu()))
    # gradients vs the CPU oracle: the solver amplifies the expf / MUFU.EX2 difference ~1000x (more on very unbalanced n : m);
    # the 1e-5 bar is held against the reference's CUDA kernels (tools/fuzz_parity4.py) and between the two GPU paths here
    if e[0] > 1e-5 or max(e[1:3]) > 3e-3 or max(e[3:]) > 1e-5:
        bad("EMD MISMATCH", dict(b=b, n=n, m=m), *e)
    ar = ad.clone().requires_grad_(True)
    w = torch.randn(b, generator=gen()).to(dev)
    (match_cost(ar, cd) * w).sum().backward()
    cases += 1
    if rel(ar.grad.cpu(), (f1 * w.view(b, 1, 1)).cpu()) > 1e-5:
        bad("EMD AUTOGRAD MISMATCH", dict(b=b, n=n, m=m), rel(ar.grad.cpu(), (f1 * w.view(b, 1, 1)).cpu()))

# ---- auction EMD ------------------------------------------------------------------------------------------------------
for it in range(max(3, iters // 3)):
    b, n = int(rng.integers(1, 4)), int(rng.choice([1024, 2048, 3072]))
    a, c = synthetic.s1_near(b, n)
    eps, its = float(rng.choice([0.005, 0.002, 0.01])), int(rng.choice([10, 30, 50]))
    dist, assign = emdModule()(a.to(dev), c.to(dev), eps, its)
    ed, ea, _ = oracle.auction_emd(a.numpy(), c.numpy(), eps, its)
    cases += 1
    if not np.array_equal(assign.cpu().numpy(), ea) or rel(dist.cpu(), ed) > 1e-5:
        bad("AUCTION MISMATCH", dict(b=b, n=n, eps=eps, iters=its), float((assign.cpu().numpy() != ea).mean()), rel(dist.cpu(), ed))
print(f"fuzz2 seed {seed}: {cases} cases, {fails} failures")
