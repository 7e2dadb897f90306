"""Programmatic dependent launch A/B (GPU box; PCC_B200_LIB must point at a -DPCC_PDL build, tools/build_variant.sh):
times the graphed ChamferEMD step, the fused EMD op and NNDistance for several PCC_PDL_MASK values in ONE process and
checks that loss and gradient are bit-identical to the mask-0 run."""
import hashlib
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import losses, synthetic  # noqa: E402
from pointcloudcounterfactual_b200.structural_losses import match_cost  # noqa: E402
from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import NNDistance  # noqa: E402

dev = torch.device("cuda", 0)
B, N = 32, 2048
recon, ref = (t.to(dev) for t in synthetic.s1_near(B, N))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(replay, reps=20):
    for _ in range(3):
        replay()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        replay()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


def sha(*ts):
    h = hashlib.sha1()
    for t in ts:
        h.update(t.detach().cpu().numpy().tobytes())
    return h.hexdigest()[:12]


def graph_of(fn):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    return g, out


out = {}
for mask in [int(a, 0) for a in sys.argv[1:]] or [0, 7]:
    os.environ["PCC_PDL_MASK"] = str(mask)
    step = losses.GraphedLossStep(losses.chamfer_emd, recon, ref, dev)
    r = {"step_us": round(timed(step), 2)}
    step()
    step.synchronize()
    r["step_sha"] = sha(step.loss_device, step.grad)
    g, res = graph_of(lambda: NNDistance(recon, ref))
    r["nndistance_us"] = round(timed(g.replay), 2)
    r["nn_sha"] = sha(*res)
    rq = recon.clone().requires_grad_(True)
    g, res = graph_of(lambda: torch.autograd.grad(match_cost(rq, ref).sum(), rq))
    r["emd_fwd_bwd_us"] = round(timed(g.replay), 2)
    r["emd_sha"] = sha(*res)
    out[f"mask{mask}"] = r
    print(json.dumps({f"mask{mask}": r}), flush=True)
base = out.get("mask0")
if base:
    for k, r in out.items():
        for key in ("step_sha", "nn_sha", "emd_sha"):
            assert r[key] == base[key], f"{k}: {key} differs from mask 0"
    print("bitwise identical to mask 0: ok")
