"""GPU: kNN graph construction / argKmin against the CPU oracle and the reference fixtures."""
import numpy as np
import pytest
import torch

import oracle
from conftest import rel_err
from pointcloudcounterfactual_b200 import _lib, keops, neighbour_ops, synthetic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("c,n,k,b", [
    (3, 1024, 20, 4), (3, 2048, 25, 2), (3, 2048, 4, 2), (3, 300, 16, 3), (3, 33, 33, 2), (3, 5, 1, 2),
    (3, 4096, 32, 1), (3, 3000, 25, 1), (3, 1500, 20, 2), (3, 1025, 1, 1), (3, 32, 32, 2), (3, 5000, 8, 1),
    (64, 1024, 20, 2), (64, 200, 20, 2), (128, 384, 25, 1), (17, 130, 7, 2), (256, 256, 20, 1),
    (64, 2048, 25, 2), (128, 1024, 20, 1), (128, 2048, 25, 1), (96, 1500, 25, 1), (32, 2048, 32, 1), (64, 1000, 31, 1),
])
def test_knn_indices_bit_exact_vs_oracle(cuda, c, n, k, b):
    x = synthetic.knn_xyz(b, n) if c == 3 else synthetic.knn_features(b, c, n)
    idx, dist = neighbour_ops.knn_indices(x.to(cuda), k, return_dist=True)
    eidx, edist = oracle.knn(x.numpy(), k, return_dist=True)
    assert idx.dtype == torch.int64 and idx.shape == (b, n, k)
    assert np.array_equal(idx.cpu().numpy(), eidx)
    assert np.array_equal(dist.cpu().numpy(), edist)  # same sequential fma chain => same bits


def test_feature_knn_tensor_core_path_edge_cases(cuda):
    """C % 32 == 0 routes to the tcgen05 candidate generator + exact re-rank: massive duplicates (candidate overflow ->
    exact brute force), ragged N (TMA zero fill), k = 32, large norms."""
    x = synthetic.knn_features(1, 64, 64).repeat(1, 1, 8)  # every point 8 times
    assert np.array_equal(neighbour_ops.knn(x.to(cuda), 20).cpu().numpy(), oracle.knn(x.numpy(), 20))
    x = synthetic.knn_features(1, 64, 64).repeat(1, 1, 16)  # n = 1024: second-generation kernel, list overflow
    assert np.array_equal(neighbour_ops.knn(x.to(cuda), 20).cpu().numpy(), oracle.knn(x.numpy(), 20))
    for (b, c, n, k) in [(2, 64, 777, 20), (1, 96, 530, 32), (1, 32, 1111, 4), (3, 128, 900, 8)]:
        x = synthetic.knn_features(b, c, n) * 7.5 + 1.0
        idx, dist = neighbour_ops.knn_indices(x.to(cuda), k, return_dist=True)
        eidx, edist = oracle.knn(x.numpy(), k, return_dist=True)
        assert np.array_equal(idx.cpu().numpy(), eidx) and np.array_equal(dist.cpu().numpy(), edist)


@pytest.mark.parametrize("b,c,n,k", [(2, 64, 1024, 20), (1, 128, 2048, 25), (2, 32, 600, 8)])
def test_feature_knn_structured_heavy_tailed_features(cuda, b, c, n, k):
    """What chained EdgeConv layers feed the feature kNN: a low-dimensional manifold embedded in C channels, far from the
    origin (distances tiny against the norms), plus a few outliers with huge norms.  The TF32 candidate bounds are per
    key; the result must stay bit-identical to the exact fp32 oracle."""
    g = torch.Generator().manual_seed(31 + n)
    z = torch.randn(b, 3, n, generator=g)                                   # the cloud itself
    a = torch.randn(1, c, 3, generator=g)
    x = torch.nn.functional.leaky_relu(a @ z + torch.randn(1, c, 1, generator=g) * 2.0, 0.2)  # (b,c,n), rank <= 3 + relu
    x[:, :, ::97] *= 30.0                                                   # outliers: norms ~1000x the typical one
    x[:, :, 1::211] = x[:, :, 0::211][..., :x[:, :, 1::211].shape[-1]]      # and some exact duplicates
    x = x.contiguous()
    idx, dist = neighbour_ops.knn_indices(x.to(cuda), k, return_dist=True)
    eidx, edist = oracle.knn(x.numpy(), k, return_dist=True)
    assert np.array_equal(idx.cpu().numpy(), eidx)
    assert np.array_equal(dist.cpu().numpy(), edist)


def _heavy_tailed(b, c, n, seed):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(b, 3, n, generator=g)
    a = torch.randn(1, c, 3, generator=g)
    x = torch.nn.functional.leaky_relu(a @ z + torch.randn(1, c, 1, generator=g) * 2.0, 0.2)
    x[:, :, ::97] *= 30.0
    x[:, :, 1::211] = x[:, :, 0::211][..., :x[:, :, 1::211].shape[-1]]
    return x.contiguous()


@pytest.mark.parametrize("kind,b,c,n,k", [
    ("iid", 2, 64, 1024, 20), ("iid", 2, 64, 2048, 25), ("iid", 1, 32, 2048, 32), ("iid", 1, 64, 1000, 31),
    ("iid", 2, 64, 777, 20), ("iid", 1, 32, 1111, 4), ("iid", 3, 64, 300, 16), ("iid", 1, 64, 256, 1),
    ("scaled", 2, 64, 1024, 20), ("dup16", 1, 64, 1024, 20), ("dup8", 1, 32, 512, 20), ("heavy", 2, 64, 1024, 20),
    ("heavy", 2, 32, 600, 8), ("heavy", 1, 64, 2048, 25), ("grid", 1, 64, 1024, 20), ("tiny", 1, 64, 512, 12),
])
def test_feature_knn_bf16_split_path_bit_exact(cuda, monkeypatch, kind, b, c, n, k):
    """The experimental bf16-split tcgen05 path for indices-only feature kNN (opt-in: PCC_KNN_BF=1, read per call):
    precise scores order the candidates, the exact fp32 chain is evaluated only inside the ambiguity band.  The indices
    must equal the oracle's bit for bit on iid, rescaled, duplicated (massive exact ties), heavy-tailed (outliers with
    huge norms, exact duplicates), quantised (many near-ties) and tiny-magnitude features, in both layouts."""
    if kind == "iid":
        x = synthetic.knn_features(b, c, n)
    elif kind == "scaled":
        x = synthetic.knn_features(b, c, n) * 7.5 + 1.0
    elif kind == "dup16":
        x = synthetic.knn_features(b, c, n // 16).repeat(1, 1, 16)
    elif kind == "dup8":
        x = synthetic.knn_features(b, c, n // 8).repeat(1, 1, 8)
    elif kind == "heavy":
        x = _heavy_tailed(b, c, n, 31 + n)
    elif kind == "grid":
        x = (synthetic.knn_features(b, c, n) * 2).round() / 2
    else:
        x = synthetic.knn_features(b, c, n) * 1e-4
    x = x.contiguous()
    monkeypatch.setenv("PCC_KNN_BF", "1")
    want = oracle.knn(x.numpy(), k)
    r0 = _lib.route_counts()
    got = neighbour_ops.knn(x.to(cuda), k)
    r1 = _lib.route_counts()
    assert r1["knn_bf"] == r0["knn_bf"] + 1
    assert np.array_equal(got.cpu().numpy(), want)
    # point-major, through the KeOps expression of the reference's pykeops_knn
    from pointcloudcounterfactual_b200.keops import LazyTensor
    xt = x.transpose(2, 1).contiguous().to(cuda)
    got_pm = ((LazyTensor(xt[:, :, None, :]) - LazyTensor(xt[:, None, :, :])) ** 2).sum(-1).argKmin(k, dim=2)
    assert _lib.route_counts()["knn_bf"] == r1["knn_bf"] + 1
    assert np.array_equal(got_pm.cpu().numpy(), want)


def test_knn_ties_lowest_index(cuda):
    a, _ = synthetic.s3_ties(2, 512, pool=64)  # heavy duplication: many exact ties
    x = a.transpose(1, 2).contiguous()
    idx = neighbour_ops.knn(x.to(cuda), 20)
    assert np.array_equal(idx.cpu().numpy(), oracle.knn(x.numpy(), 20))
    z = torch.zeros(1, 3, 64)
    assert torch.equal(neighbour_ops.knn(z.to(cuda), 5).cpu(), torch.arange(5).expand(1, 64, 5))
    # more exact ties than the warp-cooperative kernel's candidate buffer: the flagged cloud is redone by the
    # one-thread-per-query kernel, the other cloud of the batch is left alone
    zz = torch.cat([torch.zeros(1, 3, 1024), synthetic.knn_xyz(1, 1024)])
    got = neighbour_ops.knn(zz.to(cuda), 30).cpu()
    assert torch.equal(got[0], torch.arange(30).expand(1024, 30))
    assert np.array_equal(got[1].numpy(), oracle.knn(zz[1:].numpy(), 30)[0])
    a, _ = synthetic.s3_ties(2, 2048, pool=32)
    x = a.transpose(1, 2).contiguous()
    assert np.array_equal(neighbour_ops.knn(x.to(cuda), 25).cpu().numpy(), oracle.knn(x.numpy(), 25))


@pytest.mark.parametrize("n,k", [(1024, 20), (1000, 32), (2048, 25), (1500, 4), (3000, 20), (4096, 8)])
def test_xyz_knn_candidate_overflow_answered_in_kernel(cuda, monkeypatch, n, k):
    """More exact ties than the warp-cooperative kernel's candidate buffer (knn3w_overflow: k rounds of "next smallest
    (distance, index)" inside the kernel, teams of 1, 2 and 4 warps): a collapsed cloud, a cloud of 16 distinct points,
    a cloud with one NaN point, next to an ordinary cloud that must be left alone."""
    monkeypatch.setenv("PCC_KNN3_SIMT", "1")  # read per call: the SIMT xyz kernels for every n
    pool = synthetic.knn_xyz(1, 16)[0]                                  # (3, 16)
    few = pool[:, torch.arange(n) % 16]                                 # every point 1/16 of the cloud
    nanc = torch.zeros(3, n)
    nanc[:, 7] = float("nan")
    x = torch.stack([torch.zeros(3, n), few, synthetic.knn_xyz(1, n)[0], nanc]).contiguous()
    got = neighbour_ops.knn(x.to(cuda), k).cpu().numpy()
    want = oracle.knn(x[:3].numpy(), k)
    assert np.array_equal(got[:3], want)
    # the NaN point is never a neighbour of the others; rows of the real points: the lowest k indices except 7
    real = np.array([i for i in range(n) if i != 7][:k])
    rows = np.array([i for i in range(n) if i != 7])
    assert np.array_equal(got[3][rows], np.broadcast_to(real, (n - 1, k)))


@pytest.mark.parametrize("case", ["xyz_k20", "xyz_k4", "feat64_k20", "feat128_k25"])
def test_knn_golden_reference(cuda, golden, case):
    g = golden["knn"]
    x, k = torch.from_numpy(g[f"{case}_x"]), int(g[f"{case}_k"])
    idx = neighbour_ops.knn(x.to(cuda), k).cpu().numpy()
    assert np.array_equal(idx, g[f"{case}_keops_idx"])
    tidx = g[f"{case}_torch_idx"]  # reference torch path (GEMM form + topk): near-ties may be ordered differently
    mism = idx != tidx
    assert mism.mean() < 1e-3
    if mism.any():  # every disagreement must be a near-tie in fp64
        xd = torch.from_numpy(g[f"{case}_x"]).double()
        d = (xd[:, :, :, None] - xd[:, :, None, :]).pow(2).sum(1).numpy()
        bb, qq, tt = np.nonzero(mism)
        gap = np.abs(d[bb, qq, idx[bb, qq, tt]] - d[bb, qq, tidx[bb, qq, tt]])
        assert (gap <= 1e-5 * np.maximum(d[bb, qq, idx[bb, qq, tt]], 1e-6)).all()


def test_lazytensor_patterns(cuda):
    """The four KeOps expression patterns of the reference."""
    t1, t2 = synthetic.s2_far(2, 200, 333)
    a, c = t1.to(cuda), t2.to(cuda)
    d = neighbour_ops.square_distance(a, c)
    dense = oracle.square_distance(t1.numpy(), t2.numpy())
    order2 = np.argsort(dense, axis=2, kind="stable")
    order1 = np.argsort(dense, axis=1, kind="stable")
    assert np.array_equal(d.argKmin(5, dim=2).cpu().numpy(), order2[:, :, :5])
    assert np.array_equal(d.argmin(axis=2).cpu().numpy(), order2[:, :, :1])
    assert np.array_equal(d.argmin(axis=1).cpu().numpy(), order1[:, :1, :].transpose(0, 2, 1))
    assert d.argmin(axis=1).shape == (2, 333, 1)
    assert rel_err(d.sum(1).cpu().numpy()[..., 0], dense.sum(1)) < 1e-5
    # quantize.py:20-32 shape pattern: (B*codes, 1, D) against (B*codes, book, D)
    xq = torch.randn(64, 1, 4, generator=torch.Generator().manual_seed(1))
    book = torch.randn(64, 16, 4, generator=torch.Generator().manual_seed(2))
    dq = neighbour_ops.pykeops_square_distance(xq.to(cuda), book.to(cuda))
    exp = oracle.square_distance(xq.numpy(), book.numpy())
    assert np.array_equal(dq.argmin(axis=2).cpu().numpy()[:, 0, 0], np.argsort(exp, 2, kind="stable")[:, 0, 0])
    assert dq.sum(1).shape == (64, 16, 1)


def test_graph_ops_golden(cuda, golden):
    g = golden["graph"]
    x = torch.from_numpy(g["x"]).to(cuda)
    empty = torch.empty(0)
    idx, feat = neighbour_ops.get_graph_features(x, empty, k=8)
    assert np.array_equal(idx.cpu().numpy(), g["gf_idx"]) and np.array_equal(feat.cpu().numpy(), g["gf_feat"])
    assert np.array_equal(neighbour_ops.graph_max_pooling(x, empty, k=8).cpu().numpy(), g["gmp"])
    assert rel_err(neighbour_ops.get_local_covariance(x, empty, k=8).cpu().numpy(), g["cov"]) < 1e-5
    assert rel_err(neighbour_ops.graph_filtering(x, k=4).cpu().numpy(), g["filt"]) < 1e-5
    f = torch.from_numpy(g["f"]).to(cuda)
    idx, feat = neighbour_ops.get_graph_features(f, empty, k=6)
    assert np.array_equal(idx.cpu().numpy(), g["gf16_idx"]) and np.array_equal(feat.cpu().numpy(), g["gf16_feat"])
    # precomputed indices are passed through untouched (neighbour_ops.py:88-91)
    idx2, _ = neighbour_ops.get_graph_features(f, idx, k=6)
    assert idx2 is idx


def test_full_size_properties(cuda):
    """BASELINE config 2 (B=32, N=1024, k=20; xyz and 64-dim): self first, sorted, matches dense fp64 top-k sets."""
    for x in (synthetic.knn_xyz(32, 1024), synthetic.knn_features(32, 64, 1024)):
        xd = x.to(cuda)
        idx, dist = neighbour_ops.knn_indices(xd, 20, return_dist=True)
        assert torch.equal(idx[..., 0], torch.arange(1024, device=cuda).expand(32, -1))
        assert (dist[..., 1:] >= dist[..., :-1]).all() and (dist[..., 0] == 0).all()
        dense = (xd.double()[:, :, :, None] - xd.double()[:, :, None, :]).pow(2).sum(1)
        kth = dense.topk(20, largest=False)[0][..., -1]
        picked = torch.gather(dense, 2, idx)
        assert (picked <= kth.unsqueeze(-1) * (1 + 1e-5) + 1e-9).all()
        e = oracle.knn(x[:2].numpy(), 20)
        assert np.array_equal(idx[:2].cpu().numpy(), e)


def test_errors(cuda):
    x = torch.zeros(1, 3, 8, device=cuda)
    with pytest.raises(RuntimeError, match="invalid shape"):
        neighbour_ops.knn(x, 9)  # k > n, like torch.topk
    with pytest.raises(RuntimeError, match="unsupported"):
        neighbour_ops.knn(torch.zeros(1, 3, 300, device=cuda), 129)
    assert neighbour_ops.knn(torch.zeros(0, 3, 8, device=cuda), 2).shape == (0, 8, 2)
    idx = neighbour_ops.index_k_neighbours([np.random.default_rng(0).random((64, 3)).astype(np.float32)], 4)
    assert idx.shape == (1, 64, 4) and (idx[0, :, 0] == np.arange(64)).all()


def test_dataset_side_index_k_neighbours(cuda):
    """index_k_neighbours (neighbour_ops.py:16-24; consumer src/data/modelnet.py:150-156 stores it as int16 `index_{k}`):
    ragged list of clouds, batched by size, against the oracle and -- away from exact ties -- scikit-learn's KDTree,
    which is what the reference runs."""
    from sklearn.neighbors import KDTree

    rng = np.random.default_rng(5)
    pcs = [rng.standard_normal((n, 3)).astype(np.float32) for n in (256, 100, 256, 256, 100)]
    for chunk in (2, 2048):
        got = [neighbour_ops.index_k_neighbours([pc], 8)[0] for pc in pcs]  # one by one
        same = [p for p in pcs if p.shape[0] == 256]
        batched = neighbour_ops.index_k_neighbours(same, 8, chunk=chunk)
        assert batched.shape == (3, 256, 8) and batched.dtype == np.int64
        for row, ref in zip(batched, [g for g, p in zip(got, pcs) if p.shape[0] == 256]):
            assert np.array_equal(row, ref)
    for pc, g in zip(pcs, got):
        assert np.array_equal(g, oracle.knn(pc.T[None].copy(), 8)[0])
        assert np.array_equal(g, KDTree(pc).query(pc, 8, return_distance=False))
        assert np.array_equal(g.astype(np.short), g)  # the int16 on-disk format holds them


@pytest.mark.parametrize("b,c,n,k", [(2, 3, 300, 4), (2, 64, 1024, 20), (1, 128, 2048, 25), (3, 16, 77, 32), (1, 7, 8192, 3)])
def test_fused_graph_gather_forward_backward(cuda, b, c, n, k):
    """get_neighbours / get_graph_features run as one gather kernel (+ one scatter-add kernel backward); they must
    reproduce the reference's torch composition (neighbour_ops.py:85-94,113-119): forward bit for bit, the gradient to
    fp32 rounding (atomics change the summation order)."""
    g = torch.Generator().manual_seed(11)
    x = torch.randn(b, c, n, generator=g).to(cuda)
    idx = torch.randint(0, n, (b, n, k), generator=g).to(cuda)
    flat = idx.view(b, 1, k * n).expand(-1, c, -1)

    def composed(t, mode):
        nb = torch.gather(t, 2, flat).view(b, c, n, k)
        if not mode:
            return nb
        centre = t.unsqueeze(3).expand(-1, -1, -1, k)
        return torch.cat([nb - centre, centre], dim=1).contiguous()

    for mode, fn in ((0, lambda t: neighbour_ops.get_neighbours(t, idx, k)[1]),
                     (1, lambda t: neighbour_ops.get_graph_features(t, idx, k)[1])):
        a, r = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        out, ref = fn(a), composed(r, mode)
        assert torch.equal(out, ref)
        w = torch.randn(ref.shape, generator=g).to(cuda)
        (out * w).sum().backward()
        (ref * w).sum().backward()
        assert rel_err(a.grad.cpu().numpy(), r.grad.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("b,c,n,k,kind", [(2, 8, 2048, 25, "hub"), (2, 8, 1024, 20, "random"), (1, 4, 301, 5, "random"),
                                          (1, 4, 512, 16, "one_target"), (2, 8, 1605, 1, "random"), (1, 3, 77, 3, "random")])
def test_graph_gather_backward_is_reproducible(cuda, b, c, n, k, kind):
    """The gather backward sums over the target-sorted edge list in a fixed order (no atomics): two runs agree bit for
    bit, also when a few targets collect most edges (feature-space hubs) or one target collects all of them; shapes
    whose (point, neighbour) plane is not a multiple of 4 floats take the atomic kernel and still match float64."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(b, c, n, generator=g).to(cuda)
    if kind == "hub":
        idx = torch.randint(0, n, (b, n, k), generator=g)
        idx[:, :, : k // 2] = torch.randint(0, 4, (b, n, k // 2), generator=g)  # four hubs with in-degree ~ n*k/8
    elif kind == "one_target":
        idx = torch.full((b, n, k), 7, dtype=torch.int64)
    else:
        idx = torch.randint(0, n, (b, n, k), generator=g)
    idx = idx.to(cuda)
    w = torch.randn(b, 2 * c, n, k, generator=g).to(cuda)
    grads = []
    for _ in range(2):
        a = x.clone().requires_grad_(True)
        (neighbour_ops.get_graph_features(a, idx, k)[1] * w).sum().backward()
        grads.append(a.grad.clone())
    if (n * k) % 4 == 0:
        assert torch.equal(grads[0], grads[1])
    w64 = w.double()
    ref = torch.zeros(b, c, n, dtype=torch.float64, device=cuda)
    ref.scatter_add_(2, idx.view(b, 1, n * k).expand(-1, c, -1), w64[:, :c].reshape(b, c, n * k))
    ref += (w64[:, c:] - w64[:, :c]).sum(3)
    assert rel_err(grads[0].cpu().numpy(), ref.cpu().numpy()) < 1e-5


def test_vq_nearest_codeword_pattern_large_batch(cuda):
    """quantize.py:20-32 at training size: B * n_codes = 131072 one-query problems against a repeated 16 x 4 book
    (more problems than the 65535-cloud grid limit of the general kernels) run on the one-thread-per-query kernel."""
    g = torch.Generator().manual_seed(5)
    batch, n_codes, book, dim = 512, 256, 16, 4
    x = torch.randn(batch, n_codes * dim, generator=g)
    codebook = torch.randn(n_codes, book, dim, generator=g)
    codebook[3, 7] = codebook[3, 2]  # an exact tie: the lower index must win
    x_flat = x.view(batch * n_codes, 1, dim)
    rep = codebook.repeat(batch, 1, 1)
    d = neighbour_ops.pykeops_square_distance(x_flat.to(cuda), rep.to(cuda))
    idx = d.argmin(axis=2).cpu().numpy()[:, 0, 0]
    exp = oracle.square_distance(x_flat.numpy(), rep.numpy())
    assert np.array_equal(idx, np.argsort(exp, 2, kind="stable")[:, 0, 0])
    assert rel_err(d.sum(1).cpu().numpy()[..., 0], exp.sum(1)) < 1e-5


@pytest.mark.parametrize("b,n,k", [(3, 512, 4), (2, 2048, 4), (2, 300, 6), (1, 5000, 3)])
def test_fused_graph_filtering_forward_backward(cuda, b, n, k):
    """graph_filtering (neighbour_ops.py:122-133) as one launch per direction against the reference's torch
    composition (same kNN indices), value and gradient, including the path through sigma (a per-cloud mean)."""
    x = synthetic.knn_xyz(b, n).to(cuda)
    idx = neighbour_ops.knn(x, k)

    def composed(t):
        nb = torch.gather(t, 2, idx.view(b, 1, k * n).expand(-1, 3, -1)).view(b, 3, n, k)[..., 1:]
        diff = t.unsqueeze(-1).expand(-1, -1, -1, k - 1) - nb
        dist = torch.sqrt(abs((diff ** 2).sum(1)))
        sigma = torch.clamp(dist[..., 0:1].mean(1, keepdim=True), min=0.005)
        weights = torch.exp(-(dist / sigma))
        x_weight = weights.sum(2).unsqueeze(1).expand(-1, 3, -1)
        return (1 + x_weight) * t - (weights.unsqueeze(1).expand(-1, 3, -1, -1) * nb).sum(-1)

    a, r = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    out, ref = neighbour_ops.graph_filtering(a, k), composed(r)
    assert rel_err(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 1e-5
    w8 = torch.randn(ref.shape, generator=torch.Generator().manual_seed(3)).to(cuda)
    (out * w8).sum().backward()
    (ref * w8).sum().backward()
    assert rel_err(a.grad.cpu().numpy(), r.grad.cpu().numpy()) < 1e-4  # fp32 sums of ~N terms through sigma
    # tiny cloud scale: the 0.005 clamp is active and blocks the sigma path
    sc = 2.0 ** -10  # a power of two: the kNN lists of the scaled cloud are identical
    a2, r2 = (x * sc).clone().requires_grad_(True), (x * sc).clone().requires_grad_(True)
    o2, f2 = neighbour_ops.graph_filtering(a2, k), composed(r2)
    (o2 * w8).sum().backward()
    (f2 * w8).sum().backward()
    assert rel_err(o2.detach().cpu().numpy(), f2.detach().cpu().numpy()) < 1e-5
    assert rel_err(a2.grad.cpu().numpy(), r2.grad.cpu().numpy()) < 1e-4


@pytest.mark.parametrize("n,k,b", [(1024, 20, 4), (2048, 25, 33), (2048, 4, 2), (300, 8, 3), (1500, 20, 2), (2000, 32, 1), (257, 1, 2),
                                   (512, 16, 2)])
def test_xyz_knn_tensor_core_filter_bit_exact(cuda, monkeypatch, n, k, b):
    """knn3_tc_kernel (fp16 tcgen05 scores as the candidate filter, exact re-evaluation of the candidates): indices AND
    distances equal the oracle's bit for bit on every shape it accepts -- forced here also where the dispatcher prefers the
    SIMT kernel -- including duplicated points (massive exact ties -> the exact-scan path), both layouts."""
    monkeypatch.setenv("PCC_KNN3_TC", "1")
    x = synthetic.knn_xyz(b, n)
    x[0, :, 1::7] = x[0, :, 0::7][..., :x[0, :, 1::7].shape[-1]]   # exact duplicates in cloud 0
    if b > 1:
        x[1] = x[1, :, :1]                                          # cloud 1 collapsed to one point
    c0 = _lib_routes()
    idx, dist = neighbour_ops.knn_indices(x.to(cuda), k, return_dist=True)
    assert _lib_routes()["knn3_tc"] == c0["knn3_tc"] + 1
    eidx, edist = oracle.knn(x.numpy(), k, return_dist=True)
    assert np.array_equal(idx.cpu().numpy(), eidx) and np.array_equal(dist.cpu().numpy(), edist)
    xt = x.transpose(1, 2).contiguous().to(cuda)                    # point-major through the KeOps pattern
    assert torch.equal(keops.argkmin(xt, xt, k), idx)


def _lib_routes():
    from pointcloudcounterfactual_b200 import _lib

    return _lib.route_counts()
