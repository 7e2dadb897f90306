"""Accuracy and time of pcc_gemm_tf32x3 against torch.bmm (cuBLAS fp32 / TF32) on the EdgeConv shapes.
    python tools/gemm_time.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import edgeconv  # noqa: E402

dev = torch.device("cuda", 0)


def ev(fn, reps=30):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


B, N = 32, 2048
for c, cout in ((3, 64), (64, 64), (64, 128), (128, 256)):
    g = torch.Generator().manual_seed(c)
    x = torch.randn(B, c, N, generator=g).to(dev)
    ws = (torch.randn(2 * cout, c, generator=g) * 0.2).to(dev)
    guv = torch.randn(B, N, 2 * cout, generator=g).to(dev)
    uv = torch.empty(B, N, 2 * cout, device=dev)
    gx = torch.empty(B, c, N, device=dev)
    KS = 4
    gwb = torch.empty(B * KS, 2 * cout, c, device=dev)
    c2 = 2 * cout
    f = lambda: edgeconv.gemm_nt(x, (c * N, 1, N), ws, (0, c, 1), uv, (N * c2, c2, 1), B, N, c2, c)
    gi = lambda: edgeconv.gemm_nt(guv, (N * c2, c2, 1), ws, (0, 1, c), gx, (c * N, 1, N), B, N, c, c2)
    gw = lambda: edgeconv.gemm_nt(guv, (N * c2, 1, c2), x, (c * N, N, 1), gwb, (c2 * c, c, 1), B, c2, c, N, KS)
    wst = ws.t().unsqueeze(0).expand(B, -1, -1)
    tf = lambda: torch.bmm(x.transpose(1, 2), wst)
    tgi = lambda: torch.bmm(ws.t().unsqueeze(0).expand(B, -1, -1), guv.transpose(1, 2))
    tgw = lambda: torch.bmm(guv.transpose(1, 2), x.transpose(1, 2))
    f(); gi(); gw()
    w64 = torch.einsum("bcn,jc->bnj", x.double(), ws.double())
    e_f = ((uv.double() - w64).abs().max() / w64.abs().max()).item()
    e_t = ((tf().double() - w64).abs().max() / w64.abs().max()).item()
    gw64 = torch.einsum("bnj,bcn->bjc", guv.double(), x.double())
    e_gw = ((gwb.view(B, KS, c2, c).sum(1).double() - gw64).abs().max() / gw64.abs().max()).item()
    e_tgw = ((tgw().double() - gw64).abs().max() / gw64.abs().max()).item()
    print(f"C={c:3d} Cout={cout:3d}: tcgen05 3xTF32 fwd {ev(f):6.1f} us, dX {ev(gi):6.1f} us, dW {ev(gw):6.1f} us | cuBLAS fp32 fwd "
          f"{ev(tf):6.1f}, dX {ev(tgi):6.1f}, dW {ev(tgw):6.1f} | max rel err fwd {e_f:.1e} (cuBLAS {e_t:.1e}), dW {e_gw:.1e} (cuBLAS {e_tgw:.1e})")
