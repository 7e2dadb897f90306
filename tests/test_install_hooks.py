"""CPU: install() registers the module aliases and rebinds the reference's hot-path functions right after the
reference's own modules are imported (post-import hook) -- checked on a throw-away ``src`` package holding the
reference's call sites (tests/ref_callsites.py).  CPU tensors keep running the reference's ORIGINAL functions."""
import importlib
import sys

import pytest
import torch

import ref_callsites
from pointcloudcounterfactual_b200 import install as inst
from pointcloudcounterfactual_b200 import losses, neighbour_ops


@pytest.fixture()
def ref_pkg(tmp_path, monkeypatch):
    root = ref_callsites.write_package(tmp_path)
    monkeypatch.syspath_prepend(str(root))
    ref_callsites.purge()
    yield root
    inst.uninstall()
    ref_callsites.purge()


def test_post_import_hook_rebinds_before_importers_copy_the_names(ref_pkg):
    inst.install()
    enc = importlib.import_module("src.module.encoders")       # imports src.utils.neighbour_ops for the first time
    nops = sys.modules["src.utils.neighbour_ops"]
    for name in inst.PATCHED_NEIGHBOUR_OPS:
        assert hasattr(getattr(nops, name), "_pcc_b200_original"), name
    assert enc.get_graph_features is nops.get_graph_features    # the `from ... import` copy is the routed function
    mal = importlib.import_module("src.train.metrics_and_losses")
    for name in inst.PATCHED_LOSSES:
        assert hasattr(getattr(mal, name), "_pcc_b200_original"), name
    assert not hasattr(nops.torch_knn, "_pcc_b200_original")    # the reference's CPU path is left alone


def test_install_after_import_rebinds_existing_copies(ref_pkg):
    inst.install(patch_reference=False)                         # aliases + KeOps shim only
    enc = importlib.import_module("src.module.encoders")
    nops = sys.modules["src.utils.neighbour_ops"]
    original = nops.get_graph_features
    assert enc.get_graph_features is original and not hasattr(original, "_pcc_b200_original")
    inst.install()
    assert nops.get_graph_features._pcc_b200_original is original
    assert enc.get_graph_features is nops.get_graph_features
    inst.uninstall()
    assert nops.get_graph_features is original and enc.get_graph_features is original


def test_cpu_tensors_keep_the_reference_functions(ref_pkg):
    inst.install()
    nops = importlib.import_module("src.utils.neighbour_ops")
    mal = importlib.import_module("src.train.metrics_and_losses")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 40, generator=g)
    idx = nops.knn(x, 5)                                        # routed -> original -> torch_knn
    assert torch.equal(idx, neighbour_ops.torch_knn(x, 5))
    idx2, feat = nops.get_graph_features(x, torch.empty(0), 5)
    assert feat.shape == (2, 6, 40, 5) and torch.equal(idx2, idx)
    t1, t2 = torch.randn(2, 30, 3, generator=g), torch.randn(2, 20, 3, generator=g)
    d = torch.cdist(t1, t2) ** 2
    assert torch.allclose(mal.torch_chamfer(t1, t2), d.min(2)[0].sum(1) + d.min(1)[0].sum(1), atol=1e-4)
    with pytest.raises(RuntimeError, match="CUDA"):             # the B200 operators themselves have no CPU path
        losses.torch_chamfer(t1, t2)


def test_module_aliases_registered(ref_pkg):
    inst.install()
    import emd
    import pykeops
    import structural_losses
    from pykeops.torch import LazyTensor
    from structural_losses import match_cost, nn_distance  # noqa: F401

    assert structural_losses.__name__.startswith("pointcloudcounterfactual_b200")
    assert hasattr(emd, "emdModule") and pykeops._pcc_b200_shim and LazyTensor is not None
