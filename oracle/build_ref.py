"""Compile the REFERENCE's own CUDA extensions, unmodified, from where they lie under /root/reference into
oracle/_ref/ (git-ignored, but shipped to the GPU box with the repo snapshot).  TEST INFRASTRUCTURE ONLY.

    python oracle/build_ref.py

Outputs (pybind11 torch extensions, importable with :func:`load_ref`):
    oracle/_ref/structural/structural_losses_backend_ref.so   <- external/pytorch_structural_losses/src/*.cu,*.cpp
    oracle/_ref/emd/emd_backend_ref.so                        <- external/emd/src/emd.cpp, emd_cuda.cu
Nothing is copied from the reference: torch.utils.cpp_extension compiles the sources in place and only writes
objects into oracle/_ref/.  The reference's setup.py passes no arch flags (setup.py:13-22); sm_100a is used here
because that is what the GPU box runs.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/external")
OUT = HERE / "_ref"

_EXT = {
    "structural_losses_backend_ref": (
        "structural",
        [REF / "pytorch_structural_losses/src/approxmatch.cu",
         REF / "pytorch_structural_losses/src/nndistance.cu",
         REF / "pytorch_structural_losses/src/structural_loss.cpp"],
    ),
    "emd_backend_ref": ("emd", [REF / "emd/src/emd.cpp", REF / "emd/src/emd_cuda.cu"]),
}


def build() -> None:
    from torch.utils.cpp_extension import load

    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    for name, (sub, sources) in _EXT.items():
        bdir = OUT / sub
        bdir.mkdir(parents=True, exist_ok=True)
        if (bdir / f"{name}.so").exists():
            continue
        load(name=name, sources=[str(s) for s in sources], build_directory=str(bdir),
             extra_cuda_cflags=["-O3", "-gencode", "arch=compute_100a,code=sm_100a"], extra_cflags=["-O2", "-w"],
             is_python_module=False, verbose=False)


def available(name: str) -> bool:
    sub = _EXT[name][0]
    return (OUT / sub / f"{name}.so").exists()


def load_ref(name: str):
    """Import a prebuilt reference extension from oracle/_ref (never rebuilds: /root/reference is absent on the GPU box)."""
    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    sub = _EXT[name][0]
    path = OUT / sub / f"{name}.so"
    if not path.exists():
        raise FileNotFoundError(f"{path} not built; run `python oracle/build_ref.py` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    if not REF.exists():
        print("reference sources not present; nothing built", file=sys.stderr)
        sys.exit(0)
    build()
    for n in _EXT:
        print(n, "ok" if available(n) else "MISSING")
