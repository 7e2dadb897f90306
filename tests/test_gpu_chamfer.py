"""GPU: Chamfer nn_distance forward/backward (through the C ABI) against the CPU oracle and the reference fixtures."""
import numpy as np
import pytest
import torch

import oracle
from conftest import rel_err
from pointcloudcounterfactual_b200 import losses, synthetic
from pointcloudcounterfactual_b200.structural_losses import nn_distance
from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import NNDistance, NNDistanceGrad

pytestmark = pytest.mark.gpu
TOL = 1e-5  # north_star: fp32 relative tolerance for distances and gradients; indices are bit-exact


def _run_fwd(a, c, dev):
    d1, i1, d2, i2 = NNDistance(a.to(dev), c.to(dev))
    return d1.cpu().numpy(), i1.cpu().numpy(), d2.cpu().numpy(), i2.cpu().numpy()


@pytest.mark.parametrize("maker,b,n,m", [
    ("s1", 4, 512, 512), ("s2", 3, 300, 777), ("s2", 2, 2049, 130), ("s3", 3, 640, 640),
    ("s2", 2, 1, 5), ("s2", 2, 7, 1), ("s2", 1, 9, 4100), ("s1", 2, 2048, 2048),
])
def test_forward_bit_exact_vs_oracle(cuda, maker, b, n, m):
    if maker == "s1":
        a, c = synthetic.s1_near(b, n)
    elif maker == "s3":
        a, c = synthetic.s3_ties(b, n, pool=max(8, n // 4))
    else:
        a, c = synthetic.s2_far(b, max(n, 2), max(m, 2))
        a, c = a[:, :n].contiguous(), c[:, :m].contiguous()
    got = _run_fwd(a, c, cuda)
    exp = oracle.nn_distance(a.numpy(), c.numpy())
    assert np.array_equal(got[1], exp[1]) and np.array_equal(got[3], exp[3])  # indices: lowest index on ties
    assert np.array_equal(got[0], exp[0]) and np.array_equal(got[2], exp[2])  # same fma order => same bits


def test_forward_golden_reference_values(cuda, golden):
    g = golden["chamfer"]
    for case in ("s1", "s2", "s3"):
        a, c = torch.from_numpy(g[f"{case}_t1"]), torch.from_numpy(g[f"{case}_t2"])
        d1, i1, d2, i2 = _run_fwd(a, c, cuda)
        assert np.array_equal(i1, g[f"{case}_idx_axis2"]) and np.array_equal(i2, g[f"{case}_idx_axis1"])
        ta, tc = a.to(cuda).requires_grad_(True), c.to(cuda).requires_grad_(True)
        loss = losses.pykeops_chamfer(ta, tc)
        loss.sum().backward()
        assert rel_err(loss.detach().cpu().numpy(), g[f"{case}_keops_loss"]) < TOL
        assert rel_err(ta.grad.cpu().numpy(), g[f"{case}_keops_g1"]) < TOL
        assert rel_err(tc.grad.cpu().numpy(), g[f"{case}_keops_g2"]) < TOL
        assert rel_err(losses.torch_chamfer(ta, tc).detach().cpu().numpy(), g[f"{case}_torch_loss"]) < TOL


@pytest.mark.parametrize("b,n,m", [(3, 512, 512), (2, 300, 777), (2, 2048, 2048), (1, 5, 1)])
def test_backward_vs_oracle_and_deterministic(cuda, b, n, m):
    a, c = synthetic.s2_far(b, max(n, 2), max(m, 2))
    a, c = a[:, :n].contiguous(), c[:, :m].contiguous()
    g = torch.Generator().manual_seed(7)
    gd1, gd2 = torch.randn(b, n, generator=g), torch.randn(b, m, generator=g)
    _, i1, _, i2 = oracle.nn_distance(a.numpy(), c.numpy())
    e1, e2 = oracle.nn_distance_grad(a.numpy(), c.numpy(), i1, i2, gd1.numpy(), gd2.numpy())
    args = (a.to(cuda), c.to(cuda), torch.from_numpy(i1).to(cuda), torch.from_numpy(i2).to(cuda), gd1.to(cuda), gd2.to(cuda))
    g1, g2 = NNDistanceGrad(*args)
    assert rel_err(g1.cpu().numpy(), e1) < TOL and rel_err(g2.cpu().numpy(), e2) < TOL
    h1, h2 = NNDistanceGrad(*args)
    assert torch.equal(g1, h1) and torch.equal(g2, h2)  # bitwise reproducible (no float atomics)


def test_backward_collapsed_cloud_heavy_lists(cuda):
    """Early in training the reconstruction collapses: thousands of points share one nearest neighbour."""
    b, n = 2, 1536
    ref = synthetic.s2_far(b, n)[0]
    recon = torch.zeros(b, n, 3) + 0.001 * torch.randn(b, n, 3, generator=torch.Generator().manual_seed(3))
    _, i1, _, i2 = oracle.nn_distance(recon.numpy(), ref.numpy())
    assert np.bincount(i2[0]).max() > 64  # exercises the cooperative heavy-list path
    gd1, gd2 = torch.ones(b, n) / n, torch.ones(b, n) / n
    e1, e2 = oracle.nn_distance_grad(recon.numpy(), ref.numpy(), i1, i2, gd1.numpy(), gd2.numpy())
    g1, g2 = NNDistanceGrad(recon.to(cuda), ref.to(cuda), torch.from_numpy(i1).to(cuda), torch.from_numpy(i2).to(cuda),
                            gd1.to(cuda), gd2.to(cuda))
    assert rel_err(g1.cpu().numpy(), e1) < TOL and rel_err(g2.cpu().numpy(), e2) < TOL


def test_autograd_operator_surface(cuda):
    a, c = synthetic.s1_near(2, 256)
    ta, tc = a.to(cuda).requires_grad_(True), c.to(cuda).requires_grad_(True)
    d1, d2 = nn_distance(ta, tc)
    assert d1.shape == (2, 256) and d2.shape == (2, 256) and d1.dtype == torch.float32
    (d1.sum() + 2 * d2.sum()).backward()
    e = oracle.nn_distance(a.numpy(), c.numpy())
    e1, e2 = oracle.nn_distance_grad(a.numpy(), c.numpy(), e[1], e[3], np.ones((2, 256), np.float32),
                                     2 * np.ones((2, 256), np.float32))
    assert rel_err(ta.grad.cpu().numpy(), e1) < TOL and rel_err(tc.grad.cpu().numpy(), e2) < TOL
    with pytest.raises(RuntimeError):
        NNDistance(ta.detach()[:, ::2], tc.detach())  # non-contiguous input, as the reference's CHECK_INPUT


def test_full_size_properties(cuda):
    """BASELINE size (B=32, N=M=2048): properties that need no CPU oracle."""
    a, c = synthetic.s1_near(32, 2048)
    ta, tc = a.to(cuda), c.to(cuda)
    d1, i1, d2, i2 = NNDistance(ta, tc)
    near = torch.gather(tc, 1, i1.long().unsqueeze(-1).expand(-1, -1, 3))
    dx, dy, dz = (near - ta).unbind(-1)
    recomputed = torch.addcmul(torch.addcmul(dy * dy, dx, dx), dz, dz)  # not fused: allow 1 ulp
    assert torch.allclose(d1, recomputed, rtol=3e-7, atol=0)
    dense = torch.cdist(ta.double(), tc.double()).pow(2)
    assert (d1.double() <= dense.min(2)[0] * (1 + 1e-5) + 1e-12).all()
    assert (d2.double() <= dense.min(1)[0] * (1 + 1e-5) + 1e-12).all()
    s1, j1, s2, j2 = NNDistance(ta, ta)  # self-distance: zero, index = first occurrence (here: itself)
    assert (s1 == 0).all() and torch.equal(j1, torch.arange(2048, device=cuda, dtype=torch.int32).expand(32, -1))
    # sample of clouds against the oracle
    e = oracle.nn_distance(a[:3].numpy(), c[:3].numpy())
    assert np.array_equal(i1[:3].cpu().numpy(), e[1]) and np.array_equal(d2[:3].cpu().numpy(), e[2])


def test_empty_batch(cuda):
    d1, i1, d2, i2 = NNDistance(torch.zeros(0, 16, 3, device=cuda), torch.zeros(0, 8, 3, device=cuda))
    assert d1.shape == (0, 16) and i2.shape == (0, 8)


@pytest.mark.parametrize("b,n,m", [(3, 512, 512), (2, 300, 777), (2, 2048, 2048), (1, 5, 1)])
def test_fused_chamfer_losses_match_the_unfused_composition(cuda, b, n, m):
    """pykeops_chamfer / torch_chamfer run as one fused forward (+ reduction) and ONE backward launch; they must agree
    with nn_distance followed by torch reductions, value and gradients, for non-uniform upstream gradients."""
    a, c = synthetic.s2_far(b, max(n, 2), max(m, 2))
    a, c = a[:, :n].contiguous().to(cuda), c[:, :m].contiguous().to(cuda)
    w = torch.linspace(-1.0, 2.0, b, device=cuda)
    for fused, mean in ((losses.pykeops_chamfer, True), (losses.torch_chamfer, False)):
        x, y = a.clone().requires_grad_(True), c.clone().requires_grad_(True)
        loss = fused(x, y)
        (loss * w).sum().backward()
        u, v = a.clone().requires_grad_(True), c.clone().requires_grad_(True)
        d1, d2 = nn_distance(u, v)
        ref = d2.mean(1) + d1.mean(1) if mean else d1.sum(1) + d2.sum(1)
        (ref * w).sum().backward()
        assert rel_err(loss.detach().cpu().numpy(), ref.detach().cpu().numpy()) < TOL
        assert rel_err(x.grad.cpu().numpy(), u.grad.cpu().numpy()) < TOL
        assert rel_err(y.grad.cpu().numpy(), v.grad.cpu().numpy()) < TOL


@pytest.mark.parametrize("kind,b,n,m", [("s1", 32, 2048, 2048), ("s2", 4, 1500, 2048), ("s3", 4, 2048, 2048), ("s3", 3, 700, 333),
                                        ("s1", 2, 2560, 256), ("collapsed", 2, 512, 768), ("nan", 2, 600, 600),
                                        ("far_origin", 3, 1024, 1024)])
def test_tensor_core_nndistance_bit_identical_to_simt(cuda, kind, b, n, m):
    """pcc_nndistance_tc (tcgen05 candidate filter + exact resolution) returns the SAME bits as pcc_nndistance: near / far /
    massively tied / collapsed clouds, ragged sizes, NaN and inf coordinates (exact-scan path), clouds far from the origin
    (the operands are translated and scaled per cloud)."""
    from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import NNDistance, NNDistanceTC

    if kind == "s1":
        a, c = synthetic.s1_near(b, max(n, m))
    elif kind == "s2":
        a, c = synthetic.s2_far(b, n, m)
    elif kind == "s3":
        a, c = synthetic.s3_ties(b, max(n, m))
    elif kind == "collapsed":
        a, c = torch.full((b, n, 3), 0.25), torch.full((b, m, 3), 0.25)
        c[:, 5] += 0.5
    elif kind == "nan":
        a, c = synthetic.s2_far(b, n, m)
        a[0, 3] = float("nan")
        c[0, 0, 1] = float("nan")      # key 0 NaN: sticks for every query of cloud 0 (nndistance.cu:26)
        c[1, 7] = float("inf")
        a[1, 11, 2] = float("-inf")
    else:
        a, c = synthetic.s2_far(b, n, m)
        a, c = a * 3.0 + 1000.0, c * 3.0 + 1000.0
    a, c = a[:, :n].contiguous(), c[:, :m].contiguous()
    got = NNDistanceTC(a.to(cuda), c.to(cuda))
    if kind == "nan":
        # the reference's rule (oracle, nndistance.cu:26 `if (k==0 || d<best)`): a NaN distance to key 0 sticks, later NaNs
        # are skipped.  (The SIMT kernel skips a NaN at key 0 as well -- the one corner where it departs from the rule.)
        want = [torch.from_numpy(np.asarray(t)) for t in oracle.nn_distance(a.numpy(), c.numpy())]
    else:
        want = [t.cpu() for t in NNDistance(a.to(cuda), c.to(cuda))]
    for w, g in zip(want, got):
        assert torch.equal(torch.nan_to_num(w.float(), nan=-1.0), torch.nan_to_num(g.cpu().float(), nan=-1.0))
