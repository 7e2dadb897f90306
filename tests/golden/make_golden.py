"""Generate tests/golden/*.npz by running the REFERENCE's own Python on seeded inputs.

Run in the build container (needs /root/reference; never runs on the GPU box):

    python tests/golden/make_golden.py

What is executed from the reference, unmodified:
  * ``src/utils/neighbour_ops.py`` loaded by path (torch_knn, self_square_distance, torch_square_distance,
    pykeops_knn, pykeops_square_distance, get_neighbours, get_graph_features, graph_max_pooling,
    get_local_covariance, graph_filtering).
  * ``src/module/layers.py`` loaded by path (EdgeConvLayer: Conv2d 1x1 + BatchNorm2d + activation), run after the
    reference's get_graph_features and followed by the max over k exactly as ``src/module/encoders.py:49-54`` does.
  * ``pykeops_chamfer`` and ``torch_chamfer`` -- their function bodies are extracted with ``ast`` from
    ``src/train/metrics_and_losses.py`` (the module itself cannot be imported: drytorch/torcheval are absent)
    and exec'd against the imported neighbour_ops functions.
PyKeOps is an un-vendored third-party dependency (pyproject.toml:15, ``pykeops>=2.3``, not installed, JIT
source not in the tree).  It is replaced here by ``_DenseLazy``, a dense CPU emulation of the four LazyTensor
patterns the reference uses with the published KeOps semantics: ``Sum(Square(x-y))``
in fp32, arg-reductions ascending by (value, index).  Parity at the KeOps boundary itself is therefore
UNPINNED; what the fixtures pin is everything the reference's Python does around it, and the torch path.
"""
from __future__ import annotations

import ast
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(OUT.parents[1]))

from pointcloudcounterfactual_b200 import synthetic  # noqa: E402


class _DenseLazy:
    """Dense emulation of the LazyTensor expressions used by the reference (neighbour_ops.py:37-40,81;
    metrics_and_losses.py:33,36; quantize.py:28,31)."""

    def __init__(self, t: torch.Tensor):
        self.t = t

    def __sub__(self, other: "_DenseLazy") -> "_DenseLazy":
        return _DenseLazy(self.t - other.t)

    def __pow__(self, p: int) -> "_DenseLazy":
        assert p == 2
        return _DenseLazy(self.t.double() * self.t.double())  # exact product of two fp32 values

    def sum(self, dim: int = -1) -> "_DenseLazy | torch.Tensor":
        if dim in (-1, 3) and self.t.dim() == 4 and not getattr(self, "_reduced", False):
            # sequential fp32 accumulation over the feature axis with one rounding per step (fma), which is
            # what nvcc emits for KeOps' generated `acc += (x-y)*(x-y)`; emulated through float64.
            acc = torch.zeros(self.t.shape[:-1], dtype=self.t.dtype)
            for c in range(self.t.shape[-1]):
                acc = (self.t[..., c].double() + acc.double()).float()
            out = _DenseLazy(acc)
            out._reduced = True
            return out
        return self.t.sum(dim).unsqueeze(-1)  # reduction over i or j of a (B,N,M) formula -> (B,*,1)

    def _sorted_idx(self, dim: int) -> torch.Tensor:
        return torch.sort(self.t, dim=dim, stable=True)[1]

    def argKmin(self, k: int, dim: int) -> torch.Tensor:
        idx = self._sorted_idx(dim)
        return idx.narrow(dim, 0, k).contiguous() if dim == 2 else idx.narrow(dim, 0, k).transpose(1, 2).contiguous()

    def argmin(self, axis: int) -> torch.Tensor:
        idx = self._sorted_idx(axis).narrow(axis, 0, 1)
        return idx.reshape(idx.shape[0], -1, 1)


def _install_pykeops_stub() -> None:
    pk = types.ModuleType("pykeops")
    pk.set_verbose = lambda *_a, **_k: None
    pkt = types.ModuleType("pykeops.torch")
    pkt.LazyTensor = _DenseLazy
    pk.torch = pkt
    sys.modules["pykeops"] = pk
    sys.modules["pykeops.torch"] = pkt


def load_reference_neighbour_ops():
    _install_pykeops_stub()
    spec = importlib.util.spec_from_file_location("ref_neighbour_ops", REF / "src/utils/neighbour_ops.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_layers():
    spec = importlib.util.spec_from_file_location("ref_layers", REF / "src/module/layers.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_chamfers(nops):
    src = (REF / "src/train/metrics_and_losses.py").read_text()
    tree = ast.parse(src)
    ns = {"torch": torch, "pykeops_square_distance": nops.pykeops_square_distance,
          "torch_square_distance": nops.torch_square_distance}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("pykeops_chamfer", "torch_chamfer"):
            exec(compile(ast.Module([node], []), "metrics_and_losses.py", "exec"), ns)
    return ns["pykeops_chamfer"], ns["torch_chamfer"]


def _np(t):
    return t.detach().cpu().numpy()


def main() -> None:
    torch.set_num_threads(1)
    torch.manual_seed(0)
    nops = load_reference_neighbour_ops()
    pykeops_chamfer, torch_chamfer = load_reference_chamfers(nops)

    # ---- Chamfer: reference loss values + autograd gradients on S1/S2/S3, ragged N != M too -------------
    cham = {}
    cases = {
        "s1": synthetic.s1_near(3, 256),
        "s2": synthetic.s2_far(2, 192, 320),
        "s3": synthetic.s3_ties(2, 256, pool=96),
    }
    for name, (a, c) in cases.items():
        a1 = a.clone().requires_grad_(True)
        c1 = c.clone().requires_grad_(True)
        lk = pykeops_chamfer(a1, c1)
        lk.sum().backward()
        a2 = a.clone().requires_grad_(True)
        c2 = c.clone().requires_grad_(True)
        lt = torch_chamfer(a2, c2)
        lt.sum().backward()
        dense = nops.pykeops_square_distance(a, c).t  # (B,N,M) direct form
        cham.update({
            f"{name}_t1": _np(a), f"{name}_t2": _np(c),
            f"{name}_keops_loss": _np(lk), f"{name}_keops_g1": _np(a1.grad), f"{name}_keops_g2": _np(c1.grad),
            f"{name}_torch_loss": _np(lt), f"{name}_torch_g1": _np(a2.grad), f"{name}_torch_g2": _np(c2.grad),
            f"{name}_torch_sqdist": _np(nops.torch_square_distance(a, c)).astype(np.float32),
            f"{name}_idx_axis2": _np(torch.sort(dense, dim=2, stable=True)[1][:, :, 0]),
            f"{name}_idx_axis1": _np(torch.sort(dense, dim=1, stable=True)[1][:, 0, :]),
        })
    np.savez_compressed(OUT / "chamfer.npz", **cham)

    # ---- kNN: torch path (GEMM form + topk) and KeOps path (direct form), xyz and features ------------------
    knn = {}
    kcases = {
        "xyz_k20": (synthetic.knn_xyz(2, 256), 20),
        "xyz_k4": (synthetic.knn_xyz(2, 192), 4),
        "feat64_k20": (synthetic.knn_features(2, 64, 256), 20),
        "feat128_k25": (synthetic.knn_features(1, 128, 160), 25),
    }
    for name, (x, k) in kcases.items():
        knn[f"{name}_x"] = _np(x)
        knn[f"{name}_k"] = np.int64(k)
        knn[f"{name}_torch_idx"] = _np(nops.torch_knn(x, k))
        knn[f"{name}_keops_idx"] = _np(nops.pykeops_knn(x, k))
        knn[f"{name}_self_sqdist"] = _np(nops.self_square_distance(x)).astype(np.float32)
    np.savez_compressed(OUT / "knn.npz", **knn)

    # ---- graph ops built on kNN (the callers of the path) -----------------------------------
    graph = {}
    x = synthetic.knn_xyz(2, 128)
    empty = torch.empty(0)
    # monkey-patch knn to the KeOps-semantics path so CPU execution follows what the GPU run does
    nops_knn = nops.knn
    nops.knn = nops.pykeops_knn
    idx, feat = nops.get_graph_features(x, empty, k=8)
    graph.update(x=_np(x), gf_idx=_np(idx), gf_feat=_np(feat))
    graph["gmp"] = _np(nops.graph_max_pooling(x, empty, k=8))
    graph["cov"] = _np(nops.get_local_covariance(x, empty, k=8))
    graph["filt"] = _np(nops.graph_filtering(x, k=4))
    f = synthetic.knn_features(2, 16, 128)
    idx, feat = nops.get_graph_features(f, empty, k=6)
    graph.update(f=_np(f), gf16_idx=_np(idx), gf16_feat=_np(feat))
    nops.knn = nops_knn
    np.savez_compressed(OUT / "graph.npz", **graph)

    # ---- EdgeConv layer: get_graph_features -> EdgeConvLayer -> max over k, forward and backward ----
    layers = load_reference_layers()
    nops.knn = nops.pykeops_knn
    ec = {}
    ecases = {
        # name: (input (B,C,N), k, Cout, activation factory or None)
        "xyz": (synthetic.knn_xyz(2, 128), 8, 16, None),                                   # first DGCNN layer: no act
        "feat": (synthetic.knn_features(2, 16, 128), 6, 32, lambda: torch.nn.LeakyReLU(0.2)),
    }
    for name, (x0, k, cout, act_cls) in ecases.items():
        torch.manual_seed(11)
        layer = layers.EdgeConvLayer(2 * x0.shape[1], cout, act_cls=act_cls)
        with torch.no_grad():  # non-trivial affine parameters, some channels with NEGATIVE gamma (max becomes min)
            layer.bn.weight.copy_(torch.linspace(-0.7, 1.3, cout))
            layer.bn.bias.copy_(torch.linspace(0.4, -0.4, cout))
        layer.train()
        x = x0.clone().requires_grad_(True)
        idx, feat = nops.get_graph_features(x, torch.empty(0), k=k)
        out = layer(feat).max(dim=3, keepdim=False)[0]
        gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
        out.backward(gout)
        ec.update({
            f"{name}_x": _np(x0), f"{name}_k": np.int64(k), f"{name}_idx": _np(idx),
            f"{name}_weight": _np(layer.dense.weight).reshape(cout, -1), f"{name}_gamma": _np(layer.bn.weight),
            f"{name}_beta": _np(layer.bn.bias), f"{name}_slope": np.float32(-1.0 if act_cls is None else 0.2),
            f"{name}_train_out": _np(out), f"{name}_gout": _np(gout), f"{name}_gx": _np(x.grad),
            f"{name}_gw": _np(layer.dense.weight.grad).reshape(cout, -1), f"{name}_ggamma": _np(layer.bn.weight.grad),
            f"{name}_gbeta": _np(layer.bn.bias.grad), f"{name}_running_mean": _np(layer.bn.running_mean),
            f"{name}_running_var": _np(layer.bn.running_var),
        })
        layer.eval()  # running statistics as left by the one training step above
        with torch.no_grad():
            ec[f"{name}_eval_out"] = _np(layer(nops.get_graph_features(x0, idx, k=k)[1]).max(dim=3)[0])
    nops.knn = nops_knn
    np.savez_compressed(OUT / "edgeconv.npz", **ec)

    for p in sorted(OUT.glob("*.npz")):
        print(p.name, p.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
