"""Does the Chamfer chain hide under the EMD solver?  Three captured ChamferEMD steps (B=32 x 2048, forward + backward,
device-resident inputs), timed by CUDA-graph replay:
  serial      pykeops_chamfer + match_cost on one stream (what losses.chamfer_emd does)
  fork_cd     Chamfer forward (and, through autograd's stream rule, backward) on a side stream, EMD on the capture stream
  fork_emd    EMD on a HIGH-PRIORITY side stream, Chamfer on the capture stream
Run under gpurun; prints one line per variant."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import losses, synthetic  # noqa: E402
from pointcloudcounterfactual_b200.structural_losses import match_cost  # noqa: E402

dev = torch.device("cuda", 0)
recon, ref = (t.to(dev) for t in synthetic.s1_near(32, 2048))
side = torch.cuda.Stream(dev)
side_hi = torch.cuda.Stream(dev, priority=-1)


def serial(r, t):
    return losses.pykeops_chamfer(r, t) + match_cost(r, t)


def fork_cd(r, t):
    cur = torch.cuda.current_stream(dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        cd = losses.pykeops_chamfer(r, t)
    emd = match_cost(r, t)
    cur.wait_stream(side)
    return cd + emd


def fork_emd(r, t):
    cur = torch.cuda.current_stream(dev)
    side_hi.wait_stream(cur)
    with torch.cuda.stream(side_hi):
        emd = match_cost(r, t)
    cd = losses.pykeops_chamfer(r, t)
    cur.wait_stream(side_hi)
    return cd + emd


def chamfer_only(r, t):
    return losses.pykeops_chamfer(r, t)


def emd_only(r, t):
    return match_cost(r, t)


for name, fn in (("serial", serial), ("fork_cd", fork_cd), ("fork_emd", fork_emd), ("chamfer_only", chamfer_only),
                 ("emd_only", emd_only)):
    step = losses.GraphedLossStep(fn, recon, ref, dev)
    for _ in range(5):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(30):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:13s} {e0.elapsed_time(e1) / 30 * 1e3:8.1f} us per step   loss[0] = {float(step.loss_device[0]):.6f}", flush=True)
