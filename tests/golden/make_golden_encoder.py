"""Generate tests/golden/dgcnn.npz by running the REFERENCE's own DGCNN encoder (src/module/encoders.py:31-59) on a
seeded input, forward and backward, in training mode.

Run in the build container (needs /root/reference; never runs on the GPU box):

    python tests/golden/make_golden_encoder.py

Executed from the reference, unmodified: ``src/module/encoders.py`` (class DGCNN), ``src/module/layers.py``
(EdgeConvLayer, PointsConvLayer), ``src/utils/neighbour_ops.py`` (get_graph_features, knn).  What is stubbed, because the
packages are not installed here: ``src.config`` (hydra / pydantic) -- replaced by a namespace carrying the four values the
encoder reads (n_neighbors, conv_dims, w_dim, act_cls = the reference's DEFAULT_ACT, LeakyReLU(0.2), config/torch.py:11)
-- ``src.data`` (IN_CHAN = 3, data/__init__.py:9) and PyKeOps (the dense emulation of make_golden.py).
The fixture pins the CHAINED path: four EdgeConv layers with the kNN graph recomputed in feature space before each,
concatenation, final_conv, max over the points.
"""
from __future__ import annotations

import functools
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(OUT))
import make_golden as mg  # noqa: E402

REF = mg.REF
B, N, K, W_DIM = 2, 256, 8, 32


def load_reference_encoders(nops, layers):
    act = functools.partial(torch.nn.LeakyReLU, negative_slope=0.2)
    enc_cfg = types.SimpleNamespace(n_neighbors=K, conv_dims=(64, 64, 128, 256), act_cls=act)
    cfg = types.SimpleNamespace(autoencoder=types.SimpleNamespace(model=types.SimpleNamespace(encoder=enc_cfg, w_dim=W_DIM)))

    class Experiment:
        @staticmethod
        def get_config():
            return cfg

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    src = mod("src")
    src.__path__ = []
    mod("src.config", ActClass=object, Experiment=Experiment).__path__ = []
    mod("src.config.options", Encoders=types.SimpleNamespace(DGCNN="DGCNN", LDGCNN="LDGCNN"))
    mod("src.data", IN_CHAN=3)
    mod("src.module").__path__ = []
    mod("src.utils").__path__ = []
    sys.modules["src.module.layers"] = layers
    sys.modules["src.utils.neighbour_ops"] = nops
    spec = importlib.util.spec_from_file_location("ref_encoders", REF / "src/module/encoders.py")
    m = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(m)
    except Exception as exc:  # anything below the class definitions (factory code needing the real config enums)
        if not hasattr(m, "DGCNN"):
            raise
        print("note: module body stopped after the class definitions:", type(exc).__name__, exc)
    return m


def main() -> None:
    torch.set_num_threads(1)
    nops = mg.load_reference_neighbour_ops()
    nops.knn = nops.pykeops_knn  # the CUDA branch of the reference's dispatcher (neighbour_ops.py:63-68)
    layers = mg.load_reference_layers()
    enc_mod = load_reference_encoders(nops, layers)
    torch.manual_seed(3)
    net = enc_mod.DGCNN()
    net.train()
    from pointcloudcounterfactual_b200 import synthetic
    x0 = synthetic.knn_xyz(B, N).transpose(2, 1).contiguous()  # (B,N,3): what the encoder is fed
    x = x0.clone().requires_grad_(True)
    # record the graph every layer builds (the encoder's module-level name: encoders.py:13 imported it by value)
    recorded = []
    genuine_ggf = enc_mod.get_graph_features

    def recording_ggf(t, indices, k=20):
        idx, feat = genuine_ggf(t, indices=indices, k=k)
        recorded.append(idx.clone())
        return idx, feat

    enc_mod.get_graph_features = recording_ggf
    out = net(x, torch.empty(0))
    enc_mod.get_graph_features = genuine_ggf
    assert len(recorded) == 4
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(9))
    out.backward(gout)
    fx = {"x": mg._np(x0), "k": np.int64(K), "out": mg._np(out), "gout": mg._np(gout), "gx": mg._np(x.grad)}
    for i, idx in enumerate(recorded):
        fx[f"idx{i}"] = mg._np(idx).astype(np.int16)  # the kNN graph of layer i (N = 256 fits int16)
    for name, p in net.named_parameters():
        fx["param." + name] = mg._np(p)
        fx["grad." + name] = mg._np(p.grad)
    for name, bf in net.named_buffers():
        fx["buffer." + name] = mg._np(bf)  # running statistics AFTER the one training step
    np.savez_compressed(OUT / "dgcnn.npz", **fx)
    print("dgcnn.npz", (OUT / "dgcnn.npz").stat().st_size, "bytes; out", tuple(out.shape), "params", len(list(net.parameters())))


if __name__ == "__main__":
    main()
