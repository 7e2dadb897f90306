"""Candidate statistics of the bf16-split feature kNN (needs the KBF_STATS variant:
   tools/build_variant.sh kbfstats "-DKBF_STATS"; PCC_B200_LIB=pointcloudcounterfactual_b200/_lib/variants/kbfstats.so)."""
import ctypes, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pointcloudcounterfactual_b200 import _lib as L, neighbour_ops, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
lib = L.load()
lib.pcc_knn_bf_stats.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong * 8)()
for (b, c, n, k) in ((32, 64, 1024, 20), (32, 64, 2048, 25), (32, 32, 1024, 20), (32, 64, 2048, 4)):
    x = synthetic.knn_features(b, c, n).to(dev)
    lib.pcc_knn_bf_stats(buf, 1)
    neighbour_ops.knn(x, k)
    lib.pcc_knn_bf_stats(buf, 1)
    q = max(1, buf[0])
    print(f"b={b} c={c} n={n} k={k}: queries {buf[0]}, exact scans {buf[1]}, candidates mean {buf[2] / q:.1f} max {buf[3]}, "
          f"ambiguous (exact evals) mean {buf[4] / q:.2f}, list overflows {buf[5]}")
