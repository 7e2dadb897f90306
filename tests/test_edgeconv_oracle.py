"""CPU: the EdgeConv oracle (oracle/edgeconv_ref.py) against the fixture produced by the reference's own layers.py /
neighbour_ops.py (tests/golden/edgeconv.npz), forward, backward, running statistics, train and eval mode -- and the
algebra the CUDA kernels rest on (W.[x_j-x_i; x_i] = W1 x_j + (W2-W1) x_i; max through a monotone map)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import edgeconv_ref

G = np.load(Path(__file__).parent / "golden" / "edgeconv.npz")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def case(name):
    t = {k[len(name) + 1:]: torch.from_numpy(np.asarray(G[k])) for k in G.files if k.startswith(name + "_")}
    t["slope"] = None if float(t["slope"]) < 0 else float(t["slope"])
    return t


@pytest.mark.parametrize("name", ["xyz", "feat"])
def test_oracle_matches_reference_layer(name):
    c = case(name)
    x = c["x"].double().requires_grad_(True)
    w = c["weight"].double().requires_grad_(True)
    gamma = c["gamma"].double().requires_grad_(True)
    beta = c["beta"].double().requires_grad_(True)
    cout = w.shape[0]
    out, rm, rv = edgeconv_ref.edge_conv_max(x, c["idx"], w, gamma, beta, torch.zeros(cout), torch.ones(cout), True, 0.1,
                                             1e-5, c["slope"])
    out.backward(c["gout"].double())
    assert rel(out.detach(), c["train_out"]) < 1e-5
    assert rel(rm, c["running_mean"]) < 1e-5 and rel(rv, c["running_var"]) < 1e-5
    assert rel(x.grad, c["gx"]) < 2e-5 and rel(w.grad, c["gw"]) < 2e-5
    assert rel(gamma.grad, c["ggamma"]) < 2e-5 and rel(beta.grad, c["gbeta"]) < 2e-5
    ev, _, _ = edgeconv_ref.edge_conv_max(c["x"], c["idx"], c["weight"], c["gamma"], c["beta"], c["running_mean"],
                                          c["running_var"], False, 0.1, 1e-5, c["slope"])
    assert rel(ev, c["eval_out"]) < 1e-5


@pytest.mark.parametrize("name", ["xyz", "feat"])
def test_point_decomposition_is_the_same_layer(name):
    """u = W1 x, v = (W2 - W1) x; per channel the extremum over k (max for gamma >= 0, min otherwise) of u_j + v_i, then
    the affine map and the activation: what pcc_edgeconv_forward computes, in float64."""
    c = case(name)
    x, idx, w = c["x"].double(), c["idx"], c["weight"].double()
    b, ch, n = x.shape
    k = idx.shape[2]
    cout = w.shape[0]
    u = torch.einsum("oc,bcn->bno", w[:, :ch], x)
    v = torch.einsum("oc,bcn->bno", w[:, ch:] - w[:, :ch], x)
    y = torch.gather(u, 1, idx.reshape(b, n * k, 1).expand(-1, -1, cout)).view(b, n, k, cout) + v.unsqueeze(2)
    mean, var = y.mean(dim=(0, 1, 2)), y.var(dim=(0, 1, 2), unbiased=False)
    gamma, beta = c["gamma"].double(), c["beta"].double()
    ext = torch.where(gamma >= 0, y.max(dim=2)[0], y.min(dim=2)[0])
    z = (ext - mean) / torch.sqrt(var + 1e-5) * gamma + beta
    if c["slope"] is not None:
        z = torch.where(z > 0, z, z * c["slope"])
    assert rel(z.transpose(1, 2), c["train_out"]) < 1e-5
