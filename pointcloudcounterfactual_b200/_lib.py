"""ctypes loader for libpcc_b200.so (the C ABI declared in include/pcc_b200.h).

There is NO CPU fallback: if the shared library is missing, or CUDA is not available when an operator is called,
a RuntimeError is raised.  torch is used only for device memory, streams and autograd plumbing.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
# PCC_B200_LIB: alternative build of the same library (kernel A/B experiments on the GPU box)
LIB_PATH = Path(os.environ["PCC_B200_LIB"]) if os.environ.get("PCC_B200_LIB") else _PKG / "_lib" / "libpcc_b200.so"
_lib: ctypes.CDLL | None = None

_vp = ctypes.c_void_p
_i = ctypes.c_int

_SIGNATURES = {
    # name: (restype, argtypes)
    "pcc_version": (ctypes.c_char_p, []),
    "pcc_status_string": (ctypes.c_char_p, [_i]),
    "pcc_launch_count": (ctypes.c_uint64, []),
    "pcc_route_names": (ctypes.c_char_p, []),
    "pcc_route_count": (ctypes.c_int64, [ctypes.c_char_p]),
    "pcc_nndistance": (_i, [_i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pcc_nndistance_tc": (_i, [_i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pcc_nndistancegrad": (_i, [_i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pcc_chamfer_reduce": (_i, [_i, _i, _vp, _i, _vp, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pcc_chamfer_reduce_grad": (_i, [_i, _i, _vp, _i, _vp, _vp, _vp, _vp, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp]),
    "pcc_approxmatch": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pcc_matchcost": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pcc_matchcostgrad": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pcc_matchcost_fused": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pcc_approxmatch_sweep": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, ctypes.c_float, _i, _vp]),
    "pcc_knn": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pcc_argkmin": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pcc_graph_gather": (_i, [_i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "pcc_graph_gather_grad": (_i, [_i, _i, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "pcc_graph_edge_sort_bytes": (ctypes.c_longlong, [_i, _i, _i]),
    "pcc_graph_edge_sort": (_i, [_i, _i, _i, _vp, _vp, _vp]),
    "pcc_graph_gather_grad_presorted": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pcc_edgeconv_forward": (_i, [_i, _i, _i, _i] + [_vp] * 6 + [_i, ctypes.c_float, ctypes.c_float, _i, ctypes.c_float]
                             + [_vp] * 7),
    "pcc_edgeconv_backward": (_i, [_i, _i, _i, _i] + [_vp] * 6 + [_i, _i, ctypes.c_float] + [_vp] * 8),
    "pcc_gemm_tf32x3": (_i, [_i, _i, _i, _i, _i, _vp] + [ctypes.c_longlong] * 3 + [_vp] + [ctypes.c_longlong] * 3 + [_vp]
                        + [ctypes.c_longlong] * 3 + [_vp]),
    "pcc_graph_filtering": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pcc_graph_filtering_grad": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pcc_emd_forward": (_i, [_i, _i, _i] + [_vp] * 14 + [ctypes.c_float, _i, _vp]),
    "pcc_emd_backward": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load() -> ctypes.CDLL:
    """Load the CUDA library; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m pointcloudcounterfactual_b200.build` "
                "(there is no CPU or PyTorch fallback for these operators)"
            )
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def launch_count() -> int:
    return int(load().pcc_launch_count())


def route_counts() -> dict[str, int]:
    """Per kernel family, how often the library dispatched to it (pcc_route_count): evidence of WHICH kernel ran."""
    lib = load()
    return {n: int(lib.pcc_route_count(n.encode())) for n in lib.pcc_route_names().decode().split(",")}


def check(status: int, what: str, ok: int = 0) -> None:
    if status != ok:
        msg = load().pcc_status_string(status).decode()
        raise RuntimeError(f"{what} failed: {msg} (status {status})")


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def stream_of(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def require_cuda(*tensors: torch.Tensor, contiguous: bool = True, dtype: torch.dtype | None = torch.float32) -> None:
    """Same input contract as the reference binding's CHECK_INPUT (structural_loss.cpp:6-8), plus dtype/device."""
    dev = None
    for k, t in enumerate(tensors):
        if not t.is_cuda:
            raise RuntimeError(f"argument {k} must be a CUDA tensor (no CPU fallback in pointcloudcounterfactual_b200)")
        if contiguous and not t.is_contiguous():
            raise RuntimeError(f"argument {k} must be contiguous")
        if dtype is not None and t.dtype != dtype:
            raise RuntimeError(f"argument {k} must be {dtype}, got {t.dtype}")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("all arguments must live on the same CUDA device")
