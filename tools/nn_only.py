"""NNDistance forward only (B=32 x 2048), timing under graph replay; results are not checked (ablation variants)."""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import synthetic
from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import NNDistance
dev = torch.device("cuda", 0)
a, c = (t.to(dev) for t in synthetic.s1_near(32, 2048))
NNDistance(a, c); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    NNDistance(a, c)
for _ in range(3): g.replay()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(50): g.replay()
e1.record(); torch.cuda.synchronize()
print(json.dumps({"nndistance_us": round(e0.elapsed_time(e1) / 50 * 1e3, 1)}))
