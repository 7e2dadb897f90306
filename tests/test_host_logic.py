"""CPU: host-side logic that needs no GPU -- shard arithmetic, the KeOps shim's expression handling, input
validation mirroring the reference's error behaviour, module aliasing for drop-in use."""
import sys

import pytest
import torch

from pointcloudcounterfactual_b200 import keops, losses, neighbour_ops, sharding
from pointcloudcounterfactual_b200.emd import emdModule
from pointcloudcounterfactual_b200.structural_losses import structural_losses_backend as slb


def test_shard_bounds_cover_batch_without_overlap():
    for batch in (0, 1, 7, 32, 33, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(8, 2, 2)


def test_shard_batch_views():
    x = torch.arange(10).view(10, 1)
    parts = [sharding.shard_batch([x], 4, r)[0] for r in range(4)]
    assert torch.equal(torch.cat(parts), x)


def test_keops_shim_builds_the_reference_expression_on_cpu():
    t1, t2 = torch.zeros(2, 5, 3), torch.zeros(2, 7, 3)
    d = ((keops.LazyTensor(t1[:, :, None, :]) - keops.LazyTensor(t2[:, None, :, :])) ** 2).sum(-1)
    assert isinstance(d, keops.SquareDistance) and d.ti.shape == (2, 5, 3) and d.tj.shape == (2, 7, 3)
    d2 = neighbour_ops.pykeops_square_distance(t1, t2)
    assert d2.ti.shape == (2, 5, 3)
    with pytest.raises(RuntimeError, match="CUDA"):  # no CPU fallback
        d.argKmin(2, dim=2)
    with pytest.raises(NotImplementedError):
        (keops.LazyTensor(t1[:, :, None, :]) - keops.LazyTensor(t2[:, None, :, :])) ** 3


def test_ops_refuse_cpu_tensors():
    a = torch.zeros(1, 8, 3)
    for fn in (lambda: slb.NNDistance(a, a), lambda: slb.ApproxMatch(a, a), lambda: neighbour_ops.knn(a.transpose(1, 2), 2),
               lambda: losses.pykeops_chamfer(a, a)):
        with pytest.raises(RuntimeError, match="CUDA"):
            fn()


def test_emd_module_shape_rules_match_reference():
    # external/emd/emd/emd_module.py:23-30 -- raised before anything touches the GPU
    m = emdModule()
    with pytest.raises(ValueError, match="same number of points"):
        m(torch.zeros(1, 1024, 3), torch.zeros(1, 2048, 3), 0.005, 5)
    with pytest.raises(ValueError, match="Batch size must be the same"):
        m(torch.zeros(1, 1024, 3), torch.zeros(2, 1024, 3), 0.005, 5)
    with pytest.raises(ValueError, match="multiple of 1024"):
        m(torch.zeros(1, 1000, 3), torch.zeros(1, 1000, 3), 0.005, 5)
    with pytest.raises(ValueError, match="should not exceed 512"):
        m(torch.zeros(513, 1024, 3), torch.zeros(513, 1024, 3), 0.005, 5)


def test_dense_torch_helpers_keep_reference_values(golden):
    g = golden["knn"]
    x = torch.from_numpy(g["xyz_k20_x"])
    ref = torch.from_numpy(g["xyz_k20_self_sqdist"])
    assert torch.allclose(neighbour_ops.self_square_distance(x), ref, atol=2e-6)
    c = golden["chamfer"]
    t1, t2 = torch.from_numpy(c["s2_t1"]), torch.from_numpy(c["s2_t2"])
    assert torch.allclose(neighbour_ops.torch_square_distance(t1, t2), torch.from_numpy(c["s2_torch_sqdist"]), atol=2e-6)
    same = (neighbour_ops.torch_knn(x, 20).numpy() == g["xyz_k20_torch_idx"]).mean()
    assert same > 0.999


def test_install_registers_reference_module_names():
    from pointcloudcounterfactual_b200 import install

    saved = {k: sys.modules.get(k) for k in ("structural_losses", "emd", "emd_backend", "pykeops", "pykeops.torch")}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        install.install()
        import emd_backend
        import pykeops
        import structural_losses
        from emd import emdModule as E
        from pykeops.torch import LazyTensor
        from structural_losses.structural_losses_backend import ApproxMatch, MatchCost, MatchCostGrad, NNDistance, NNDistanceGrad

        assert callable(structural_losses.match_cost) and callable(structural_losses.nn_distance)
        assert E is emdModule and LazyTensor is keops.LazyTensor
        assert all(callable(f) for f in (ApproxMatch, MatchCost, MatchCostGrad, NNDistance, NNDistanceGrad))
        assert callable(emd_backend.forward) and callable(emd_backend.backward)
        pykeops.set_verbose(False)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
