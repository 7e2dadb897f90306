"""Quick GPU check of the tcgen05 feature-kNN path against the exact SIMT path / oracle."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import oracle  # noqa: E402
from pointcloudcounterfactual_b200 import neighbour_ops, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
for (b, c, n, k) in [(1, 32, 256, 8), (2, 64, 256, 20), (2, 64, 1024, 20), (1, 128, 384, 25), (2, 64, 300, 16), (1, 256, 512, 20)]:
    x = synthetic.knn_features(b, c, n)
    t0 = time.time()
    idx, dist = neighbour_ops.knn_indices(x.to(dev), k, return_dist=True)
    torch.cuda.synchronize()
    e, ed = oracle.knn(x.numpy(), k, return_dist=True)
    ok = np.array_equal(idx.cpu().numpy(), e)
    okd = np.array_equal(dist.cpu().numpy(), ed)
    print(f"b={b} c={c} n={n} k={k}: idx_equal={ok} dist_equal={okd} mismatch_frac={(idx.cpu().numpy() != e).mean():.5f} ({time.time()-t0:.2f}s)", flush=True)
# duplicated points: overflow path
x = synthetic.knn_features(1, 64, 64).repeat(1, 1, 8)
idx = neighbour_ops.knn(x.to(dev), 20)
print("ties:", np.array_equal(idx.cpu().numpy(), oracle.knn(x.numpy(), 20)))
def graph_us(fn, reps=20):
    fn(); torch.cuda.synchronize()
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


for (b, c, n, k) in [(32, 64, 1024, 20), (32, 128, 1024, 20), (32, 64, 2048, 25), (8, 64, 1024, 20)]:
    x = synthetic.knn_features(b, c, n).to(dev)
    print(f"graph replay B={b} C={c} N={n} k={k}: {graph_us(lambda: neighbour_ops.knn(x, k)):.1f} us", flush=True)
