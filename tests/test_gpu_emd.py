"""GPU: approximate-matching EMD (ApproxMatch / MatchCost / MatchCostGrad / fused match_cost) and the auction EMD
against the CPU oracle."""
import numpy as np
import pytest
import torch

import oracle
from conftest import rel_err
from pointcloudcounterfactual_b200 import losses, synthetic
from pointcloudcounterfactual_b200.emd import emdModule
from pointcloudcounterfactual_b200.structural_losses import match_cost
from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import (
    ApproxMatch, MatchCost, MatchCostFused, MatchCostGrad)

pytestmark = pytest.mark.gpu
TOL = 1e-5  # north_star: EMD costs and gradients within 1e-5 relative
# The approxmatch iteration amplifies rounding-level differences in exp() about 1000x into `match` and the gradients
# (measured on the CPU: replacing expf(x) by exp2f(x*log2e) in the oracle moves match by 2e-4 absolute and the
# gradients by 1e-4 relative, while the cost moves by 1e-7).  The CPU oracle cannot reproduce MUFU.EX2's bits, so
# against the ORACLE match/gradients are held to AMP_TOL; the 1e-5 bar is enforced against the reference's own CUDA
# kernels (tests/test_ref_cuda_parity.py), whose arithmetic the sweep kernel reproduces operation by operation, and
# between the materialised and fused GPU paths below.
AMP_TOL = 1e-3


def _clouds(kind, b, n, m):
    if kind == "s1":
        return synthetic.s1_near(b, n)
    a, c = synthetic.s2_far(b, max(n, 2), max(m, 2))  # normalise() needs two points; slice afterwards
    return a[:, :n].contiguous(), c[:, :m].contiguous()


@pytest.mark.parametrize("kind,b,n,m", [("s1", 2, 512, 512), ("s2", 2, 384, 384), ("s2", 2, 256, 128), ("s2", 1, 100, 333),
                                        ("s1", 1, 2048, 2048), ("s2", 2, 1, 7)])
def test_approxmatch_chain_vs_oracle(cuda, kind, b, n, m):
    a, c = _clouds(kind, b, n, m)
    ematch, _ = oracle.approxmatch(a.numpy(), c.numpy())
    ecost = oracle.matchcost(a.numpy(), c.numpy(), ematch)
    eg1, eg2 = oracle.matchcostgrad(a.numpy(), c.numpy(), ematch)
    ta, tc = a.to(cuda), c.to(cuda)
    match, temp = ApproxMatch(ta, tc)
    assert match.shape == (b, m, n) and temp.shape == (b, 2 * (n + m))
    # element-wise: absolute tolerance relative to the row scale (entries span 30 orders of magnitude)
    assert np.abs(match.cpu().numpy() - ematch).max() < AMP_TOL * max(1.0, float(ematch.max()))
    assert rel_err(match.sum(1).cpu().numpy(), ematch.sum(1)) < 1e-4
    cost = MatchCost(ta, tc, match)
    assert rel_err(cost.cpu().numpy(), ecost) < TOL
    g1, g2 = MatchCostGrad(ta, tc, match)
    assert rel_err(g1.cpu().numpy(), eg1) < AMP_TOL and rel_err(g2.cpu().numpy(), eg2) < AMP_TOL
    # the same operators fed the ORACLE's match: isolates MatchCost / MatchCostGrad from ApproxMatch
    om = torch.from_numpy(ematch).to(cuda)
    assert rel_err(MatchCost(ta, tc, om).cpu().numpy(), ecost) < TOL
    h1, h2 = MatchCostGrad(ta, tc, om)
    assert rel_err(h1.cpu().numpy(), eg1) < TOL and rel_err(h2.cpu().numpy(), eg2) < TOL
    # fused path (no match matrix)
    fc, f1, f2 = MatchCostFused(ta, tc)
    assert rel_err(fc.cpu().numpy(), ecost) < TOL
    assert rel_err(f1.cpu().numpy(), eg1) < AMP_TOL and rel_err(f2.cpu().numpy(), eg2) < AMP_TOL
    # fused vs materialised on the GPU share the solver, so they must agree to the fp32 bar
    assert rel_err(fc.cpu().numpy(), cost.cpu().numpy()) < TOL
    assert rel_err(f1.cpu().numpy(), g1.cpu().numpy()) < TOL and rel_err(f2.cpu().numpy(), g2.cpu().numpy()) < TOL


def test_match_cost_autograd_surface(cuda):
    a, c = synthetic.s1_near(3, 256)
    ta, tc = a.to(cuda).requires_grad_(True), c.to(cuda)
    cost = match_cost(ta, tc)
    assert cost.shape == (3,)
    w = torch.tensor([1.0, -2.0, 0.5], device=cuda)
    (cost * w).sum().backward()
    ematch, _ = oracle.approxmatch(a.numpy(), c.numpy())
    eg1, _ = oracle.matchcostgrad(a.numpy(), c.numpy(), ematch)
    assert rel_err(ta.grad.cpu().numpy(), eg1 * w.cpu().numpy()[:, None, None]) < AMP_TOL
    assert tc.grad is None
    tb = c.to(cuda).requires_grad_(True)
    match_cost(a.to(cuda), tb).sum().backward()
    _, eg2 = oracle.matchcostgrad(a.numpy(), c.numpy(), ematch)
    assert rel_err(tb.grad.cpu().numpy(), eg2) < AMP_TOL


def test_invariants_full_size(cuda):
    """BASELINE config 3 size (B=32, n=m=2048): SURVEY section 4 invariants, no CPU oracle needed."""
    a, c = synthetic.s1_near(32, 2048)
    ta, tc = a.to(cuda), c.to(cuda)
    match, _ = ApproxMatch(ta, tc)
    rows, cols = match.sum(1), match.sum(2)
    assert rows.min() > 0.99 and rows.max() <= 1 + 1e-5 and cols.min() > 0.99 and cols.max() <= 1 + 1e-5
    cost = MatchCost(ta, tc, match)
    fc, f1, _ = MatchCostFused(ta, tc, want_grad2=False)
    assert rel_err(fc.cpu().numpy(), cost.cpu().numpy()) < TOL
    g1, _ = MatchCostGrad(ta, tc, match)
    assert rel_err(f1.cpu().numpy(), g1.cpu().numpy()) < TOL
    same, _, _ = MatchCostFused(ta, ta, want_grad1=False, want_grad2=False)
    assert (same < 0.03 * cost).all()  # match_cost(x, x) ~ 0: the first level leaves a little mass on close neighbours
    del match
    # two clouds against the oracle at full size
    em, _ = oracle.approxmatch(a[:2].numpy(), c[:2].numpy())
    assert rel_err(fc[:2].cpu().numpy(), oracle.matchcost(a[:2].numpy(), c[:2].numpy(), em)) < TOL
    eg1, _ = oracle.matchcostgrad(a[:2].numpy(), c[:2].numpy(), em)
    assert rel_err(f1[:2].cpu().numpy(), eg1) < AMP_TOL


def test_fused_is_deterministic(cuda):
    a, c = synthetic.s2_far(4, 700, 650)
    x = MatchCostFused(a.to(cuda), c.to(cuda))
    y = MatchCostFused(a.to(cuda), c.to(cuda))
    assert all(torch.equal(p, q) for p, q in zip(x, y))


def _cull_cases():
    a, c = synthetic.s1_near(3, 2048)
    nan = a.clone()
    nan[1, 17, 2] = float("nan")
    nan[2, 5, 0] = float("inf")
    return {
        "s1": synthetic.s1_near(8, 2048), "s2": synthetic.s2_far(4, 2048, 2048), "s3_ties": synthetic.s3_ties(4, 2048),
        "ragged": synthetic.s2_far(3, 1500, 700), "ragged_multi_tile": synthetic.s2_far(2, 100, 3000),
        "two_tiles": synthetic.s1_near(2, 4096), "collapsed": (a * 1e-3, c), "scaled_x8": (a * 8, c * 8),
        "nan_inf": (nan, c),
    }


@pytest.mark.parametrize("case", ["s1", "s2", "s3_ties", "ragged", "ragged_multi_tile", "two_tiles", "collapsed",
                                  "scaled_x8", "nan_inf"])
def test_culled_sweeps_bit_identical_to_full_sweeps(cuda, case, monkeypatch):
    """Exact-zero culling (csrc/approxmatch.cu): the sweeps of the two steepest levels skip partners whose exponential
    underflows to +0.  Nothing may change: every output bit equals the full sweeps' (PCC_AM_NOCULL=1 is read per call)."""
    a, c = _cull_cases()[case]
    a, c = a.to(cuda).contiguous(), c.to(cuda).contiguous()

    def run():
        out = list(MatchCostFused(a, c, True, True))
        if a.shape[0] * a.shape[1] * c.shape[1] <= 4 * 2048 * 2048:
            out += list(ApproxMatch(a, c))[:1]   # the materialised plan; `temp` is scratch
        torch.cuda.synchronize()
        return out

    monkeypatch.setenv("PCC_AM_NOCULL", "1")
    full = run()
    monkeypatch.setenv("PCC_AM_NOCULL", "0")
    culled = run()
    for f, g in zip(full, culled):
        assert torch.equal(f.view(torch.int32), g.view(torch.int32))


@pytest.mark.parametrize("b,n,eps,iters", [(3, 1024, 0.005, 50), (2, 2048, 0.005, 50), (1, 1024, 0.002, 3000), (2, 1024, 0.01, 1),
                                           (2, 3072, 0.005, 30),   # exactly 48 KiB of dynamic shared memory + static words
                                           (1, 4096, 0.01, 20), (1, 5120, 0.005, 10)])
def test_auction_vs_oracle(cuda, b, n, eps, iters):
    a, c = synthetic.auction_clouds(b, n)
    edist, easg, _ = oracle.auction_emd(a.numpy(), c.numpy(), eps, iters)
    ta = a.to(cuda).requires_grad_(True)
    dist, asg = emdModule()(ta, c.to(cuda), eps, iters)
    assert asg.dtype == torch.int32 and dist.shape == (b, n)
    assert np.array_equal(asg.cpu().numpy(), easg)  # deterministic rounds => identical assignment
    assert rel_err(dist.detach().cpu().numpy(), edist) < TOL
    w = torch.randn(b, n, generator=torch.Generator().manual_seed(5))
    (dist * w.to(cuda)).sum().backward()
    eg = oracle.auction_emd_grad(a.numpy(), c.numpy(), w.numpy(), easg)
    assert rel_err(ta.grad.cpu().numpy(), eg) < TOL


def test_graphed_loss_step_equals_eager(cuda):
    """losses.GraphedLossStep: host -> device copies, chamfer_emd forward/backward and the loss read-back captured as one
    CUDA graph; refilling the pinned buffers and replaying gives exactly the eager results (same kernels, same order)."""
    recon, ref = synthetic.s1_near(3, 512)
    rh, th = recon.clone().pin_memory(), ref.clone().pin_memory()
    step = losses.GraphedLossStep(losses.chamfer_emd, rh, th, cuda)
    assert step.kernels_per_replay > 20
    for seed_shift in (0, 7):
        a, c = synthetic.s1_near(3, 512, first=seed_shift)
        rh.copy_(a)
        th.copy_(c)
        loss_h, grad = step()
        step.synchronize()
        r = a.to(cuda).requires_grad_(True)
        want = losses.chamfer_emd(r, c.to(cuda))
        (gw,) = torch.autograd.grad(want.sum(), r)
        assert torch.equal(loss_h, want.detach().cpu())
        assert torch.equal(grad.cpu(), gw.cpu())
    with pytest.raises(RuntimeError):
        losses.GraphedLossStep(losses.chamfer_emd, recon, ref, cuda)  # not pinned


def test_graphed_loss_step_device_resident(cuda):
    a, c = synthetic.s1_near(2, 384)
    step = losses.GraphedLossStep(losses.chamfer_emd, a.to(cuda), c.to(cuda), cuda)
    a2, c2 = synthetic.s1_near(2, 384, first=3)
    step.recon.data.copy_(a2.to(cuda))
    step.ref.copy_(c2.to(cuda))
    loss_h, grad = step()
    step.synchronize()
    r = a2.to(cuda).requires_grad_(True)
    want = losses.chamfer_emd(r, c2.to(cuda))
    (gw,) = torch.autograd.grad(want.sum(), r)
    assert torch.equal(loss_h, want.detach().cpu()) and torch.equal(grad.cpu(), gw.cpu())


def test_pipelined_loss_step_equals_eager(cuda):
    """losses.PipelinedLossStep: two staging buffers / two graphs alternate, the next step's H2D overlaps the current
    step; every step's loss and gradient equal the eager ones, and the losses come back one step late, in order."""
    b, n = 2, 384
    rh = torch.empty(b, n, 3).pin_memory()
    th = torch.empty(b, n, 3).pin_memory()
    batches = [synthetic.s1_near(b, n, first=3 * i) for i in range(5)]
    rh.copy_(batches[0][0]), th.copy_(batches[0][1])
    pipe = losses.PipelinedLossStep(losses.chamfer_emd, rh, th, cuda)
    pipe.prefetch()
    got_losses, got_grads = [], []
    for i in range(5):
        pipe.wait_prefetch()  # the prefetch of batch i has read the pinned buffers (no device-wide sync): refill for i+1
        if i + 1 < 5:
            rh.copy_(batches[i + 1][0]), th.copy_(batches[i + 1][1])
        prev = pipe.step()
        got_grads.append(pipe.grad.clone())
        if prev is not None:
            got_losses.append(prev.clone())
    got_losses.append(pipe.drain().clone())
    for (a, c), lo, gr in zip(batches, got_losses, got_grads):
        r = a.to(cuda).requires_grad_(True)
        want = losses.chamfer_emd(r, c.to(cuda))
        (gw,) = torch.autograd.grad(want.sum(), r)
        assert torch.equal(lo, want.detach().cpu()) and torch.equal(gr.cpu(), gw.cpu())
