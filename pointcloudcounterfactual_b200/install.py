"""Make the reference's scripts pick up the B200 operators without editing them.

``install()`` registers, under the module names the reference imports:
    structural_losses, structural_losses.structural_losses_backend      (metrics_and_losses.py:10)
    emd, emd_backend                                                    (external/emd/emd/emd_module.py:9)
    pykeops, pykeops.torch                                              (neighbour_ops.py:5,11)
so that ``train_autoencoder.py`` / ``train_w_autoencoder.py`` / ``generate.py`` run unchanged after
``import pointcloudcounterfactual_b200.install as i; i.install()`` (e.g. from sitecustomize).
"""
from __future__ import annotations

import sys

from . import emd as _emd
from . import keops as _keops
from . import structural_losses as _sl
from .emd import emd_backend as _emd_backend
from .structural_losses import structural_losses_backend as _slb


def install(keops: bool = True) -> None:
    sys.modules["structural_losses"] = _sl
    sys.modules["structural_losses.structural_losses_backend"] = _slb
    sys.modules["structural_losses.match_cost"] = sys.modules[_sl.__name__ + ".match_cost"]
    sys.modules["structural_losses.nn_distance"] = sys.modules[_sl.__name__ + ".nn_distance"]
    sys.modules["emd"] = _emd
    sys.modules["emd_backend"] = _emd_backend
    sys.modules["emd.emd_backend"] = _emd_backend
    if keops:
        _keops.install()
