"""kNN probe (GPU box): timings under CUDA-graph replay for the BASELINE / repo-default shapes."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import neighbour_ops, synthetic  # noqa: E402

dev = torch.device("cuda", 0)


def ev(fn, reps=30, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def graph(fn, reps=30):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return ev(g.replay, reps)


out = {}
cases = [("xyz", 32, 3, 1024, 20), ("xyz", 32, 3, 2048, 25), ("xyz", 32, 3, 2048, 4), ("xyz", 4, 3, 2048, 25),
         ("xyz", 256, 3, 2048, 4), ("feat", 32, 64, 1024, 20), ("feat", 32, 64, 2048, 25), ("feat", 32, 128, 2048, 25)]
only = sys.argv[1] if len(sys.argv) > 1 else ""
for kind, b, c, n, k in cases:
    if only and kind != only:
        continue
    x = (synthetic.knn_xyz(b, n) if c == 3 else synthetic.knn_features(b, c, n)).to(dev)
    out[f"{kind}_b{b}_c{c}_n{n}_k{k}_us"] = round(graph(lambda: neighbour_ops.knn(x, k)), 1)
print(json.dumps(out))
