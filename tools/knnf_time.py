"""Times the feature-space kNN (tcgen05 path) for the bench / DGCNN shapes (CUDA-graph replay, CUDA events)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import neighbour_ops, synthetic  # noqa: E402
sys.path.insert(0, str(Path(__file__).resolve().parent))
from knn_time import ev  # noqa: E402

dev = torch.device("cuda", 0)
for (b, c, n, k) in ((32, 64, 1024, 20), (32, 64, 2048, 25), (32, 128, 2048, 25), (32, 32, 1024, 20)):
    x = synthetic.knn_features(b, c, n).to(dev)
    print(f"feature kNN b={b} c={c} n={n} k={k}: {ev(lambda: neighbour_ops.knn(x, k)):7.1f} us", flush=True)
