"""GPU: the reference's UNCHANGED call sites (tests/ref_callsites.py restates them) reach the fast kernels after
install() -- indices bit-exact, and the library's dispatch counters (pcc_route_count) prove WHICH kernel family ran:
  * through the KeOps shim alone (the reference's own function bodies run): point-major self kNN -> knn3w (N = 1024) / knn3_tc (N = 2048) / knn_tc2,
    the two argmin reductions of pykeops_chamfer -> ONE nn_sym launch;
  * through the post-import hook: the fused operators (one-launch graph gather, graph_filtering, fused Chamfer loss).
Sizes are BASELINE.json's: B=32, N=1024 / 2048, C=3 / 64."""
import importlib
import sys

import numpy as np
import pytest
import torch

import oracle
import ref_callsites
from conftest import rel_err
from pointcloudcounterfactual_b200 import _lib, losses, neighbour_ops, synthetic
from pointcloudcounterfactual_b200 import install as inst

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ref_pkg(tmp_path, monkeypatch):
    root = ref_callsites.write_package(tmp_path)
    monkeypatch.syspath_prepend(str(root))
    ref_callsites.purge()
    yield root
    inst.uninstall()
    ref_callsites.purge()


def _delta(before):
    after = _lib.route_counts()
    return {k: after[k] - before[k] for k in after if after[k] != before[k]}


@pytest.mark.parametrize("patched", [False, True])
def test_unchanged_knn_call_sites_reach_fast_kernels(cuda, ref_pkg, patched):
    inst.install(patch_reference=patched)
    nops = importlib.import_module("src.utils.neighbour_ops")
    x3 = synthetic.knn_xyz(32, 1024).to(cuda)                     # (B,3,N) channels-first, as the encoders hold it
    xf = synthetic.knn_features(32, 64, 1024).to(cuda)
    x25 = synthetic.knn_xyz(32, 2048).to(cuda)
    for x, k, fam in ((x3, 20, "knn3w"), (xf, 20, "knn_tc2"), (x25, 25, "knn3_tc"), (x25, 4, "knn3_tc")):
        c0 = _lib.route_counts()
        idx = nops.knn(x, k)                                       # the reference's dispatcher -> pykeops_knn
        d = _delta(c0)
        assert d.get(fam, 0) == 1 and "knn_simt" not in d, (fam, d)
        if not patched:                                            # the reference's own body: transpose + LazyTensor on (x, x)
            assert d.get("pm_self", 0) == 1, d
        assert idx.dtype == torch.int64 and idx.shape == (32, x.shape[2], k)
        assert torch.equal(idx, neighbour_ops.knn(x, k))           # bit-exact with the direct channels-first entry
        assert np.array_equal(idx[:2].cpu().numpy(), oracle.knn(x[:2].cpu().numpy(), k))


@pytest.mark.parametrize("patched", [False, True])
def test_unchanged_chamfer_call_site_is_one_fused_launch(cuda, ref_pkg, patched):
    inst.install(patch_reference=patched)
    mal = importlib.import_module("src.train.metrics_and_losses")
    recon, ref = synthetic.s1_near(32, 2048)
    r = recon.to(cuda).requires_grad_(True)
    t = ref.to(cuda)
    c0 = _lib.route_counts()
    loss = mal.pykeops_chamfer(r, t)                               # two argmin reductions of ONE symbolic expression
    d = _delta(c0)
    assert d.get("nn_sym", 0) + d.get("nn_tc", 0) == 1 and "knn_simt" not in d, d
    loss.sum().backward()
    d1, i1, d2, i2 = oracle.nn_distance(recon[:4].numpy(), ref[:4].numpy())
    assert rel_err(loss[:4].detach().cpu().numpy(), d1.mean(1) + d2.mean(1)) < 1e-5
    r2 = recon.to(cuda).requires_grad_(True)
    direct = losses.pykeops_chamfer(r2, t)
    direct.sum().backward()
    assert rel_err(loss.detach().cpu().numpy(), direct.detach().cpu().numpy()) < 1e-6
    assert rel_err(r.grad.cpu().numpy(), r2.grad.cpu().numpy()) < 1e-5
    # the indices the shim hands back are the reference kernel's (lowest index on ties), (B,M,1) / (B,N,1) int64
    nops = sys.modules["src.utils.neighbour_ops"]
    dist = nops.pykeops_square_distance(r.detach(), t)
    a1, a2 = dist.argmin(axis=1), dist.argmin(axis=2)
    assert a1.shape == (32, 2048, 1) and a1.dtype == torch.int64
    assert np.array_equal(a2[:4, :, 0].cpu().numpy(), i1) and np.array_equal(a1[:4, :, 0].cpu().numpy(), i2)


def test_shim_argmin_ties_and_ragged(cuda, ref_pkg):
    inst.install(patch_reference=False)
    nops = importlib.import_module("src.utils.neighbour_ops")
    a, b = synthetic.s3_ties(3, 700)
    b = b[:, :333].contiguous()
    dist = nops.pykeops_square_distance(a.to(cuda), b.to(cuda))
    d1, i1, d2, i2 = oracle.nn_distance(a.numpy(), b.numpy())
    assert np.array_equal(dist.argmin(axis=2)[..., 0].cpu().numpy(), i1)
    assert np.array_equal(dist.argmin(axis=1)[..., 0].cpu().numpy(), i2)
    assert np.array_equal(dist.min(axis=2)[..., 0].cpu().numpy(), d1)


def test_hook_routes_graph_ops_to_fused_kernels(cuda, ref_pkg):
    inst.install()
    enc = importlib.import_module("src.module.encoders")
    nops = sys.modules["src.utils.neighbour_ops"]
    xf = synthetic.knn_features(4, 64, 2048).to(cuda)
    n0 = _lib.launch_count()
    idx, feat = enc.edge_features(xf, 25)                          # `from src.utils.neighbour_ops import get_graph_features`
    assert _lib.launch_count() - n0 <= 4                           # kNN (prep + tcgen05) + ONE gather launch
    original = nops.get_graph_features._pcc_b200_original
    idx_o, feat_o = original(xf, idx, 25)                          # the reference's torch composition on the same indices
    assert torch.equal(feat, feat_o)
    x3 = synthetic.knn_xyz(4, 2048).to(cuda)
    out = nops.graph_filtering(x3)
    ref_out = nops.graph_filtering._pcc_b200_original(x3)
    assert rel_err(out.cpu().numpy(), ref_out.cpu().numpy()) < 1e-5
    pooled = nops.graph_max_pooling(xf, idx, 25)
    assert torch.equal(pooled, nops.graph_max_pooling._pcc_b200_original(xf, idx, 25))


def test_graph_max_pooling_ignores_tf32_setting(cuda):
    x = synthetic.knn_features(2, 64, 512).to(cuda)
    idx = neighbour_ops.knn(x, 16)
    want = torch.gather(x, 2, idx.view(2, 1, -1).expand(-1, 64, -1)).view(2, 64, 512, 16).max(-1)[0]
    was = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        xg = x.clone().requires_grad_(True)
        got = neighbour_ops.graph_max_pooling(xg, idx, 16)
        assert torch.equal(got, want)
        got.sum().backward()
        assert float(xg.grad.sum()) == pytest.approx(2 * 64 * 512, rel=1e-6)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = was
