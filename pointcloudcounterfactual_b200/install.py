"""Make the reference's scripts pick up the B200 operators without editing them.

``install()`` does two things:

1.  registers, under the module names the reference imports,
        structural_losses, structural_losses.structural_losses_backend      (metrics_and_losses.py:10)
        emd, emd_backend                                                    (external/emd/emd/emd_module.py:9)
        pykeops, pykeops.torch                                              (neighbour_ops.py:5,11)
2.  rebinds, as soon as the reference's own modules are imported (a ``sys.meta_path`` hook; at once if they already
    are), the functions of the geometry hot path to the fused sm_100a operators:
        src.utils.neighbour_ops.{knn, pykeops_knn, get_neighbours, get_graph_features, graph_max_pooling, graph_filtering}
        src.train.metrics_and_losses.{pykeops_chamfer, torch_chamfer}
    A rebound function runs the B200 operator for CUDA tensors and the reference's ORIGINAL function for anything else
    (the reference's CPU path stays the reference's code -- this package has no CPU implementation of its own).
    Modules that did ``from src.utils.neighbour_ops import get_graph_features`` before ``install()`` ran are rebound too.

so that ``train_autoencoder.py`` / ``train_w_autoencoder.py`` / ``generate.py`` run unchanged after
``import pointcloudcounterfactual_b200.install as i; i.install()`` (e.g. from sitecustomize).  Without step 2 the
reference's own function bodies still work -- their KeOps expressions reach the same kernels through the ``keops`` shim
(self kNN and the fused two-direction argmin are recognised there) -- they just pay for the torch glue around them.
"""
from __future__ import annotations

import functools
import importlib.abc
import importlib.machinery
import sys
from types import ModuleType
from typing import Callable

from . import emd as _emd
from . import keops as _keops
from . import structural_losses as _sl
from .emd import emd_backend as _emd_backend
from .structural_losses import structural_losses_backend as _slb

NEIGHBOUR_OPS = "src.utils.neighbour_ops"
METRICS_AND_LOSSES = "src.train.metrics_and_losses"
PATCHED_NEIGHBOUR_OPS = ("knn", "pykeops_knn", "get_neighbours", "get_graph_features", "graph_max_pooling",
                         "graph_filtering")
PATCHED_LOSSES = ("pykeops_chamfer", "torch_chamfer")
_MARK = "_pcc_b200_original"


def _first_tensor(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if hasattr(a, "is_cuda"):
            return a
    return None


def _route(ours: Callable, original: Callable) -> Callable:
    """CUDA tensors -> the B200 operator; anything else -> the reference's own function, untouched."""

    @functools.wraps(original)
    def routed(*args, **kwargs):
        t = _first_tensor(args, kwargs)
        if t is not None and t.is_cuda:
            return ours(*args, **kwargs)
        return original(*args, **kwargs)

    setattr(routed, _MARK, original)
    return routed


def _rebind_importers(original: Callable, routed: Callable) -> None:
    """``from module import f`` copies the binding: fix the copies made before install() ran."""
    for mod in list(sys.modules.values()):
        d = getattr(mod, "__dict__", None)
        if not isinstance(d, dict):
            continue
        for name, val in list(d.items()):
            if val is original:
                d[name] = routed


def patch_module(mod: ModuleType) -> list[str]:
    """Rebind the hot-path functions of an imported reference module; returns the names it rebound."""
    from . import losses, neighbour_ops

    if mod.__name__ == NEIGHBOUR_OPS:
        src, names = neighbour_ops, PATCHED_NEIGHBOUR_OPS
    elif mod.__name__ == METRICS_AND_LOSSES:
        src, names = losses, PATCHED_LOSSES
    else:
        return []
    done = []
    for name in names:
        original = getattr(mod, name, None)
        if original is None or hasattr(original, _MARK):
            continue
        routed = _route(getattr(src, name), original)
        setattr(mod, name, routed)
        _rebind_importers(original, routed)
        done.append(name)
    return done


class _PatchingLoader(importlib.abc.Loader):
    def __init__(self, inner):
        self._inner = inner

    def create_module(self, spec):
        return self._inner.create_module(spec)

    def exec_module(self, module):
        self._inner.exec_module(module)
        patch_module(module)

    def __getattr__(self, name):  # get_code, get_source, is_package ... of the real loader
        return getattr(self._inner, name)


class _PostImportFinder(importlib.abc.MetaPathFinder):
    """Finds the two reference modules with the regular machinery and patches them right after they execute -- before
    the importing module's ``from ... import`` line copies the bindings."""

    _busy = False

    def find_spec(self, fullname, path=None, target=None):
        if fullname not in (NEIGHBOUR_OPS, METRICS_AND_LOSSES) or _PostImportFinder._busy:
            return None
        _PostImportFinder._busy = True
        try:
            spec = importlib.machinery.PathFinder.find_spec(fullname, path)
        finally:
            _PostImportFinder._busy = False
        if spec is None or spec.loader is None:
            return None
        spec.loader = _PatchingLoader(spec.loader)
        return spec


def install(keops: bool = True, patch_reference: bool = True) -> None:
    sys.modules["structural_losses"] = _sl
    sys.modules["structural_losses.structural_losses_backend"] = _slb
    sys.modules["structural_losses.match_cost"] = sys.modules[_sl.__name__ + ".match_cost"]
    sys.modules["structural_losses.nn_distance"] = sys.modules[_sl.__name__ + ".nn_distance"]
    sys.modules["emd"] = _emd
    sys.modules["emd_backend"] = _emd_backend
    sys.modules["emd.emd_backend"] = _emd_backend
    if keops:
        _keops.install()
    if patch_reference:
        if not any(isinstance(f, _PostImportFinder) for f in sys.meta_path):
            sys.meta_path.insert(0, _PostImportFinder())
        for name in (NEIGHBOUR_OPS, METRICS_AND_LOSSES):
            if name in sys.modules:
                patch_module(sys.modules[name])


def uninstall() -> None:
    """Undo ``install()`` (tests): restore the reference's functions, drop the hook and the registered modules."""
    sys.meta_path[:] = [f for f in sys.meta_path if not isinstance(f, _PostImportFinder)]
    for name in (NEIGHBOUR_OPS, METRICS_AND_LOSSES):
        mod = sys.modules.get(name)
        if mod is None:
            continue
        for attr, val in list(vars(mod).items()):
            original = getattr(val, _MARK, None)
            if original is not None:
                setattr(mod, attr, original)
                _rebind_importers(val, original)
    for name in ("structural_losses", "structural_losses.structural_losses_backend", "structural_losses.match_cost",
                 "structural_losses.nn_distance", "emd", "emd_backend", "emd.emd_backend"):
        sys.modules.pop(name, None)
    if getattr(sys.modules.get("pykeops"), "_pcc_b200_shim", False):
        sys.modules.pop("pykeops", None)
        sys.modules.pop("pykeops.torch", None)
