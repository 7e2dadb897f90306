"""Build libpcc_b200.so (hand-written CUDA for sm_100a + the C ABI of include/pcc_b200.h) in-tree with nvcc.

    python -m pointcloudcounterfactual_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT_DIR = PKG / "_lib"
LIB = OUT_DIR / "libpcc_b200.so"
SOURCES = ["lib.cu", "chamfer.cu", "chamfer_tc.cu", "knn3_tc.cu", "knn.cu", "knn_tc.cu", "knn_tc2.cu", "knn_bf.cu", "approxmatch.cu", "auction.cu", "graph.cu", "edgeconv.cu", "gemm_tc.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]
# Programmatic dependent launch (csrc/common.cuh) is compiled into the Chamfer forward chain: symmetric kernel ->
# finalize -> loss reduction (measured on B200, tools/pdl_probe.py: NNDistance 58.7 -> 56.7 us, bit-identical) and the
# feature kNN's prep -> main pair (79.3 -> 78.6 us with the trigger at the end of the prep kernel); on the EMD solver
# sweeps it is a loss.  PCC_PDL_MASK=0 in the environment switches it off at run time.
PDL_FLAGS = ["-DPCC_PDL", "-DPCC_PDL_DEFAULT_MASK=1"]
EXTRA_FLAGS = {"lib.cu": PDL_FLAGS, "chamfer.cu": PDL_FLAGS, "knn_tc2.cu": PDL_FLAGS}


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libpcc_b200.so cannot be built")
    return exe


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "pcc_b200.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not _stale():
        return LIB
    OUT_DIR.mkdir(exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        o = OUT_DIR / (Path(s).stem + ".o")
        cmd = [nvcc(), *NVCC_FLAGS, *EXTRA_FLAGS.get(s, []), "-c", str(CSRC / s), "-o", str(o)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(o))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
    link = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *objs]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
