"""CPU: pin the oracle (oracle/geom_oracle.c) against fixtures produced by the reference's own Python
(tests/golden/make_golden.py) and against the algorithm-derived invariants of SURVEY.md section 4."""
import numpy as np
import pytest

import oracle
from conftest import rel_err
from pointcloudcounterfactual_b200 import synthetic

TOL = 1e-5  # north_star: distances, EMD costs and gradients within 1e-5 relative


@pytest.mark.parametrize("case", ["s1", "s2", "s3"])
def test_nn_distance_matches_reference_chamfer(golden, case):
    g = golden["chamfer"]
    a, c = g[f"{case}_t1"], g[f"{case}_t2"]
    d1, i1, d2, i2 = oracle.nn_distance(a, c)
    # indices: bit-exact against the reference's KeOps-semantics argmin (lowest index on ties; s3 has ties)
    assert np.array_equal(i1, g[f"{case}_idx_axis2"])
    assert np.array_equal(i2, g[f"{case}_idx_axis1"])
    # reference loss values: pykeops_chamfer = mean, torch_chamfer = sum (metrics_and_losses.py:35-47)
    assert rel_err(d1.mean(1) + d2.mean(1), g[f"{case}_keops_loss"]) < TOL
    assert rel_err(d1.sum(1) + d2.sum(1), g[f"{case}_torch_loss"]) < TOL
    n, m = a.shape[1], c.shape[1]
    g1, g2 = oracle.nn_distance_grad(a, c, i1, i2, np.full_like(d1, 1.0 / n), np.full_like(d2, 1.0 / m))
    assert rel_err(g1, g[f"{case}_keops_g1"]) < TOL
    assert rel_err(g2, g[f"{case}_keops_g2"]) < TOL
    g1s, g2s = oracle.nn_distance_grad(a, c, i1, i2, np.ones_like(d1), np.ones_like(d2))
    if case != "s3":  # with exact ties torch.min's autograd may route to a different (equally near) point
        assert rel_err(g1s, g[f"{case}_torch_g1"]) < 1e-4  # torch path differentiates the GEMM form
        assert rel_err(g2s, g[f"{case}_torch_g2"]) < 1e-4


def test_square_distance_matches_reference_dense(golden):
    g = golden["chamfer"]
    d = oracle.square_distance(g["s2_t1"], g["s2_t2"])
    ref = g["s2_torch_sqdist"]  # GEMM form: cancellation error ~1e-6 absolute on unit-ball clouds
    assert np.abs(d - ref).max() < 5e-6


@pytest.mark.parametrize("case", ["xyz_k20", "xyz_k4", "feat64_k20", "feat128_k25"])
def test_knn_matches_reference(golden, case):
    g = golden["knn"]
    x, k = g[f"{case}_x"], int(g[f"{case}_k"])
    idx, dist = oracle.knn(x, k, return_dist=True)
    assert np.array_equal(idx, g[f"{case}_keops_idx"])  # direct form, stable order
    # torch path (GEMM form + topk): identical sets except where fp32 rounding reorders near-ties
    same = (idx == g[f"{case}_torch_idx"]).mean()
    assert same > 0.999
    assert (np.diff(dist, axis=-1) >= 0).all()
    assert np.array_equal(idx[..., 0], np.broadcast_to(np.arange(x.shape[2]), idx.shape[:2]))  # self first


def test_knn_tie_rule():
    x = np.zeros((1, 3, 16), np.float32)  # all points identical: ascending indices expected
    idx = oracle.knn(x, 5)
    assert np.array_equal(idx[0], np.tile(np.arange(5), (16, 1)))


def test_nn_distance_invariants():
    a, c = (t.numpy() for t in synthetic.s3_ties(2, 300, pool=64))
    d1, i1, d2, i2 = oracle.nn_distance(a, c)
    dense = oracle.square_distance(a, c)  # x-first fma order: may differ by 1 ulp from the kernel order
    assert np.array_equal(i1, np.argmin(np.where(dense <= d1[..., None] * (1 + 1e-6) + 1e-12, 0, 1), axis=2))
    d1s, i1s, _, _ = oracle.nn_distance(a, a)
    assert (d1s == 0).all()
    first = np.array([[int(np.flatnonzero((a[b] == a[b, j]).all(1))[0]) for j in range(a.shape[1])] for b in range(2)])
    assert np.array_equal(i1s, first)


def test_nn_grad_finite_difference():
    rng = np.random.default_rng(0)
    a = rng.random((1, 12, 3)).astype(np.float32)
    c = rng.random((1, 9, 3)).astype(np.float32)
    w1 = rng.random((1, 12)).astype(np.float32)
    w2 = rng.random((1, 9)).astype(np.float32)

    def loss(a64, c64):
        d = ((a64[:, :, None, :] - c64[:, None, :, :]) ** 2).sum(-1)
        return (w1 * d.min(2)).sum() + (w2 * d.min(1)).sum()

    _, i1, _, i2 = oracle.nn_distance(a, c)
    g1, g2 = oracle.nn_distance_grad(a, c, i1, i2, w1, w2)
    eps = 1e-6
    num = np.zeros_like(a, dtype=np.float64)
    for j in range(12):
        for d in range(3):
            ap = a.astype(np.float64).copy(); am = ap.copy()
            ap[0, j, d] += eps; am[0, j, d] -= eps
            num[0, j, d] = (loss(ap, c.astype(np.float64)) - loss(am, c.astype(np.float64))) / (2 * eps)
    assert np.abs(num - g1).max() < 1e-4
    assert np.isfinite(g2).all()


def test_approxmatch_invariants():
    a, c = (t.numpy() for t in synthetic.s2_far(1, 512))
    match, temp = oracle.approxmatch(a, c)
    rows, cols = match.sum(1), match.sum(2)  # match is (B, m, n)
    assert rows.min() > 0.99 and rows.max() <= 1 + 1e-5
    assert cols.min() > 0.99 and cols.max() <= 1 + 1e-5
    cost = oracle.matchcost(a, c, match)
    same, _ = oracle.approxmatch(a, a)
    assert oracle.matchcost(a, a, same)[0] < 1e-3 * cost[0]
    from scipy.optimize import linear_sum_assignment

    d = np.sqrt(((a[0][:, None, :] - c[0][None, :, :]) ** 2).sum(-1))
    r, cidx = linear_sum_assignment(d)
    assert cost[0] >= d[r, cidx].sum() * 0.999  # cost of a near-feasible plan >= exact EMD


def test_approxmatch_ragged_multipliers():
    a, c = (t.numpy() for t in synthetic.s2_far(1, 256, 128))
    match, _ = oracle.approxmatch(a, c)  # n=256 >= m=128: remainR starts at n/m = 2 (approxmatch.cu:6-12)
    assert abs(match.sum() - 256) < 1.0
    assert match.sum(1).max() <= 1 + 1e-5 and match.sum(2).max() <= 2 + 1e-5


def test_matchcostgrad_finite_difference():
    a, c = (t.numpy() for t in synthetic.s2_far(1, 24, 20))
    match, _ = oracle.approxmatch(a, c)
    g1, g2 = oracle.matchcostgrad(a, c, match)
    eps = 1e-3
    for (j, d) in [(0, 0), (5, 1), (23, 2)]:
        ap, am = a.copy(), a.copy()
        ap[0, j, d] += eps; am[0, j, d] -= eps
        num = (oracle.matchcost(ap, c, match)[0] - oracle.matchcost(am, c, match)[0]) / (2 * eps)
        assert abs(num - g1[0, j, d]) < 2e-2 * max(1.0, abs(num))
    assert np.isfinite(g2).all()


def test_auction_emd_properties():
    a, c = (t.numpy() for t in synthetic.auction_clouds(2, 1024))
    dist, asg, _ = oracle.auction_emd(a, c, 0.005, 50)
    assert (asg >= 0).all() and (asg < 1024).all()
    got = ((a - np.take_along_axis(c, asg[..., None].astype(np.int64), 1)) ** 2).sum(-1)
    assert rel_err(dist, got) < TOL
    # with the README's test setting (eps=0.002, 10 000 rounds) the assignment is a permutation and near-optimal
    dist2, asg2, _ = oracle.auction_emd(a, c, 0.002, 10000)
    assert all(len(set(asg2[b])) == 1024 for b in range(2))
    from scipy.optimize import linear_sum_assignment

    d = np.sqrt(((a[0][:, None, :] - c[0][None, :, :]) ** 2).sum(-1))
    r, cidx = linear_sum_assignment(d)
    assert np.sqrt(dist2[0]).sum() <= d[r, cidx].sum() * 1.05 + 1024 * 0.002
    with pytest.raises(ValueError):
        oracle.auction_emd(a[:, :1000], c[:, :1000], 0.005, 5)


def test_torch_ref_matches_reference(golden):
    """oracle/torch_ref.py (the CPU baseline that bench.py times) reproduces the genuine reference functions."""
    import torch

    from oracle import torch_ref

    c = golden["chamfer"]
    for case in ("s1", "s2"):
        t1, t2 = torch.from_numpy(c[f"{case}_t1"]), torch.from_numpy(c[f"{case}_t2"])
        loss, g1 = torch_ref.chamfer_fwd_bwd(t1, t2)
        assert rel_err(loss.numpy(), c[f"{case}_torch_loss"]) < 1e-6
        assert rel_err(g1.numpy(), c[f"{case}_torch_g1"]) < 1e-6
    k = golden["knn"]
    x = torch.from_numpy(k["feat64_k20_x"])
    assert (torch_ref.torch_knn(x, 20).numpy() == k["feat64_k20_torch_idx"]).mean() > 0.9999


def _approxmatch_numpy(x1, x2):
    """Independent restatement of SURVEY.md Appendix A.3 (approxmatch.cu:3-182) in float64, vectorised per sweep."""
    n, m = len(x1), len(x2)
    multi_l, multi_r = (1, n // m) if n >= m else (m // n, 1)
    d2 = ((x1[:, None, :].astype(np.float64) - x2[None, :, :]) ** 2).sum(-1)  # (n, m): k indexes cloud 1, l cloud 2
    remain_l, remain_r = np.full(n, float(multi_l)), np.full(m, float(multi_r))
    match = np.zeros((m, n))
    for j in range(7, -2, -1):
        e = np.exp(-(4.0 ** j) * d2)
        ratio_l = remain_l / (1e-9 + e @ remain_r)
        sumr = (ratio_l @ e) * remain_r
        ratio_r = np.minimum(remain_r / (sumr + 1e-9), 1.0) * remain_r
        remain_r = np.maximum(0.0, remain_r - sumr)
        w = e * ratio_l[:, None] * ratio_r[None, :]
        match += w.T
        remain_l = np.maximum(0.0, remain_l - w.sum(1))
    return match, remain_l, remain_r


@pytest.mark.parametrize("n,m", [(96, 96), (128, 64), (48, 144), (100, 37)])
def test_approxmatch_matches_independent_numpy_restatement(n, m):
    """The C oracle against a float64 numpy restatement written from the pseudo-code alone: same plan (B,m,n layout,
    integer multipliers incl. the non-divisible case, level schedule, guards), same cost and gradients.  The iteration
    amplifies fp32-vs-fp64 rounding ~1000x, hence 1e-3 on the plan and 1e-5 only on the cost."""
    a, c = (t.numpy() for t in synthetic.s2_far(1, n, m))
    match, temp = oracle.approxmatch(a, c)
    want, rl, rr = _approxmatch_numpy(a[0], c[0])
    assert match.shape == (1, m, n)
    assert np.abs(match[0] - want).max() < 1e-3 * max(want.max(), 1e-3)
    assert np.abs(temp[0, :n] - rl).max() < 1e-3 and np.abs(temp[0, n:n + m] - rr).max() < 1e-3
    diff = a[0][:, None, :].astype(np.float64) - c[0][None, :, :]  # (n, m, 3)
    dist = np.sqrt((diff ** 2).sum(-1))
    cost = (match[0].astype(np.float64).T * dist).sum()
    assert rel_err(oracle.matchcost(a, c, match), np.array([cost])) < TOL
    g1, g2 = oracle.matchcostgrad(a, c, match)
    unit = diff / np.sqrt(np.maximum((diff ** 2).sum(-1), 1e-20))[..., None]
    w = match[0].astype(np.float64).T[..., None] * unit
    assert rel_err(g1[0], w.sum(1)) < TOL and rel_err(g2[0], -w.sum(0)) < TOL


def test_auction_matches_independent_python_restatement():
    """oracle.auction_emd against a plain-Python restatement of Appendix A.5 (emd_cuda.cu:94-225) with the oracle's
    deterministic tie rule (highest source index among bids within 1e-6 wins), on a small cloud pair."""
    n = 1024
    a, c = (t.numpy() for t in synthetic.auction_clouds(1, n))
    eps, iters = np.float32(0.01), 6
    x1, x2 = a[0], c[0]
    asg, inv, price = np.full(n, -1), np.full(n, -1), np.zeros(n, np.float32)
    dx = x2[None, :, :] - x1[:, None, :]
    d2 = (dx[..., 2] * dx[..., 2] + (dx[..., 0] * dx[..., 0] + dx[..., 1] * dx[..., 1])).astype(np.float32)
    root = np.sqrt(d2).astype(np.float32)
    for it in range(iters):
        last = it == iters - 1
        un = np.flatnonzero(asg == -1)
        if len(un) == 0:
            break
        v = (np.float32(3.0) - root[un]) - price[None, :]
        best_k = v.argmax(1)
        best = v[np.arange(len(un)), best_k]
        v2 = v.copy()
        v2[np.arange(len(un)), best_k] = -np.inf
        inc = (best - v2.max(1)) + eps
        top = np.full(n, -1e9, np.float32)
        np.maximum.at(top, best_k, inc)
        winner = np.full(n, -1)
        for i, k, g in zip(un, best_k, inc):  # ascending i: the highest index within 1e-6 of the top bid wins
            if abs(g - top[k]) <= 1e-6:
                winner[k] = i
        for i, k, g in zip(un, best_k, inc):
            if last or winner[k] == i:
                if not last and inv[k] != -1:
                    asg[inv[k]] = -1
                inv[k], asg[i] = i, k
                price[k] += g
    dist, got, _ = oracle.auction_emd(a, c, float(eps), iters)
    assert np.array_equal(got[0], asg)
    assert rel_err(dist[0], d2[np.arange(n), asg]) < TOL
