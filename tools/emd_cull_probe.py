"""Culled vs full approxmatch sweeps (csrc/approxmatch.cu, "exact-zero culling"): bitwise comparison of every output of
pcc_matchcost_fused and pcc_approxmatch on S1 / S2 / S3 / collapsed / ragged / NaN clouds, and graph-replay timing of the
fused forward+backward at B=32 x 2048 with and without the cull (PCC_AM_NOCULL is read at every call)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import synthetic  # noqa: E402
from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import ApproxMatch, MatchCostFused  # noqa: E402

dev = torch.device("cuda", 0)


def run(a, c, cull):
    os.environ["PCC_AM_NOCULL"] = "0" if cull else "1"
    cost, g1, g2 = MatchCostFused(a, c, True, True)
    match, temp = ApproxMatch(a, c) if a.shape[0] * a.shape[1] * c.shape[1] <= 8 * 2048 * 2048 else (None, None)
    torch.cuda.synchronize()
    return cost, g1, g2, match


def same(x, y):
    if x is None:
        return True
    return bool(torch.equal(x.view(torch.int32), y.view(torch.int32)))


cases = {}
cases["s1 32x2048"] = synthetic.s1_near(32, 2048)
cases["s2 8x2048"] = synthetic.s2_far(8, 2048, 2048)
cases["s3 8x2048"] = synthetic.s3_ties(8, 2048)
cases["s2 ragged 4x1500x700"] = synthetic.s2_far(4, 1500, 700)
cases["s2 ragged 3x100x3000"] = synthetic.s2_far(3, 100, 3000)
cases["s1 2x4096"] = synthetic.s1_near(2, 4096)
a, c = synthetic.s1_near(4, 2048)
cases["collapsed recon"] = (a * 1e-3, c)
cases["scaled x8"] = (a * 8, c * 8)
an = a.clone()
an[1, 17, 2] = float("nan")
an[2, 5, 0] = float("inf")
cases["nan/inf points"] = (an, c)
ok = True
for name, (a, c) in cases.items():
    a, c = a.to(dev).contiguous(), c.to(dev).contiguous()
    full = run(a, c, False)
    cull = run(a, c, True)
    res = [same(x, y) for x, y in zip(full, cull)]
    ok &= all(res)
    print(f"{name:28s} cost/grad1/grad2/match bit-identical: {res}", flush=True)
print("ALL BIT-IDENTICAL" if ok else "MISMATCH")


def timed(a, c, cull, reps=20):
    os.environ["PCC_AM_NOCULL"] = "0" if cull else "1"
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            MatchCostFused(a, c, True, False)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            MatchCostFused(a, c, True, False)
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.synchronize()
        e0.record(s)
        for _ in range(reps):
            g.replay()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, gen in (("s1", lambda: synthetic.s1_near(32, 2048)), ("s2", lambda: synthetic.s2_far(32, 2048, 2048))):
    a, c = gen()
    a, c = a.to(dev), c.to(dev)
    t_full = timed(a, c, False)
    t_cull = timed(a, c, True)
    print(f"{name} B=32x2048 matchcost_fused fwd+bwd: full {t_full:.1f} us, culled {t_cull:.1f} us")
    for lv in (1, 3, 4):
        os.environ["PCC_AM_CULL_LEVELS"] = str(lv)
        print(f"   culled levels = {lv}: {timed(a, c, True):.1f} us")
    del os.environ["PCC_AM_CULL_LEVELS"]
os.environ["PCC_AM_NOCULL"] = "0"
