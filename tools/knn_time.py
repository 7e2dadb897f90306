"""Times the kNN entry points (CUDA-graph replay, CUDA events) for the bench shapes; PCC_KNN3_SIMT=1 selects the SIMT xyz path.
    python tools/knn_time.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import _lib, neighbour_ops, synthetic  # noqa: E402

dev = torch.device("cuda", 0)


def ev(fn, reps=50):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


if __name__ == "__main__":
  for (b, n, k) in ((32, 1024, 20), (32, 1024, 8), (32, 2048, 25), (32, 2048, 4), (32, 512, 16), (4, 2048, 25), (256, 1024, 20)):
      x = synthetic.knn_xyz(b, n).to(dev)
      r0 = _lib.route_counts()
      us = ev(lambda: neighbour_ops.knn(x, k))
      r1 = _lib.route_counts()
      print(f"xyz kNN b={b} n={n} k={k}: {us:7.1f} us  routes {sorted(kk for kk in r1 if r1[kk] > r0[kk])}")
