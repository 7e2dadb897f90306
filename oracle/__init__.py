"""CPU oracle for the geometry hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  ``pointcloudcounterfactual_b200`` never does (tests/test_layout.py checks).

* :mod:`oracle.geom_oracle` (C, ``oracle/geom_oracle.c``) -- fp32 restatement of the reference CUDA kernels
  (``external/pytorch_structural_losses/src/*.cu``, ``external/emd/src/emd_cuda.cu``) and of the canonical
  direct-form kNN; each C function cites the reference lines it follows.
* :mod:`oracle.torch_ref` -- restatement of the reference's torch CPU path (``src/utils/neighbour_ops.py:43-74``,
  ``src/train/metrics_and_losses.py:21-47``), the "reference torch CPU path" BASELINE.json config 1 names.

Pinning: ``tests/golden/*.npz`` were produced by importing the reference's own Python from ``/root/reference``
(``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py`` checks this oracle against them.  On the GPU
box ``tests/test_ref_cuda_parity.py`` additionally checks the product against the reference's own CUDA kernels
compiled from ``/root/reference`` into ``oracle/_ref/`` (``oracle/build_ref.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libgeom_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile oracle/geom_oracle.c with gcc (see oracle/Makefile)."""
    src = _HERE / "geom_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int)
_L = ctypes.POINTER(ctypes.c_int64)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_LIB_PATH))
        i = ctypes.c_int
        L.orc_num_threads.restype = i
        L.orc_set_num_threads.argtypes = [i]
        L.orc_nn_distance.argtypes = [i, i, _F, i, _F, _F, _I, _F, _I]
        L.orc_nn_distance_grad.argtypes = [i, i, _F, i, _F, _F, _I, _F, _I, _F, _F]
        L.orc_approxmatch.argtypes = [i, i, i, _F, _F, _F, _F]
        L.orc_matchcost.argtypes = [i, i, i, _F, _F, _F, _F]
        L.orc_matchcostgrad.argtypes = [i, i, i, _F, _F, _F, _F, _F]
        L.orc_knn.argtypes = [i, i, i, i, _F, _L, _F]
        L.orc_square_distance.argtypes = [i, i, i, i, _F, _F, _F]
        L.orc_auction_emd.argtypes = [i, i, i, _F, _F, ctypes.c_float, i, _F, _I, _F]
        L.orc_auction_emd.restype = i
        L.orc_auction_emd_grad.argtypes = [i, i, _F, _F, _F, _I, _F]
        _lib = L
    return _lib


def _f(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def _i(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a), dtype=np.int32)


def _pf(a: np.ndarray):
    return a.ctypes.data_as(_F)


def _pi(a: np.ndarray):
    return a.ctypes.data_as(_I)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(t: int) -> None:
    lib().orc_set_num_threads(int(t))


def nn_distance(xyz1, xyz2):
    """(B,N,3),(B,M,3) -> dist1 (B,N) f32, idx1 (B,N) i32, dist2 (B,M), idx2 (B,M)."""
    a, c = _f(xyz1), _f(xyz2)
    b, n, _ = a.shape
    m = c.shape[1]
    d1 = np.empty((b, n), np.float32)
    i1 = np.empty((b, n), np.int32)
    d2 = np.empty((b, m), np.float32)
    i2 = np.empty((b, m), np.int32)
    lib().orc_nn_distance(b, n, _pf(a), m, _pf(c), _pf(d1), _pi(i1), _pf(d2), _pi(i2))
    return d1, i1, d2, i2


def nn_distance_grad(xyz1, xyz2, idx1, idx2, gd1, gd2):
    a, c = _f(xyz1), _f(xyz2)
    b, n, _ = a.shape
    m = c.shape[1]
    g1 = np.empty((b, n, 3), np.float32)
    g2 = np.empty((b, m, 3), np.float32)
    i1, i2, e1, e2 = _i(idx1), _i(idx2), _f(gd1), _f(gd2)
    lib().orc_nn_distance_grad(b, n, _pf(a), m, _pf(c), _pf(e1), _pi(i1), _pf(e2), _pi(i2), _pf(g1), _pf(g2))
    return g1, g2


def approxmatch(xyz1, xyz2, want_match: bool = True):
    """-> match (B,m,n) (or None), temp (B, 2(n+m))."""
    a, c = _f(xyz1), _f(xyz2)
    b, n, _ = a.shape
    m = c.shape[1]
    match = np.empty((b, m, n), np.float32) if want_match else None
    temp = np.empty((b, 2 * (n + m)), np.float32)
    lib().orc_approxmatch(b, n, m, _pf(a), _pf(c), _pf(match) if want_match else None, _pf(temp))
    return match, temp


def matchcost(xyz1, xyz2, match):
    a, c, mt = _f(xyz1), _f(xyz2), _f(match)
    b, n, _ = a.shape
    m = c.shape[1]
    out = np.empty((b,), np.float32)
    lib().orc_matchcost(b, n, m, _pf(a), _pf(c), _pf(mt), _pf(out))
    return out


def matchcostgrad(xyz1, xyz2, match):
    a, c, mt = _f(xyz1), _f(xyz2), _f(match)
    b, n, _ = a.shape
    m = c.shape[1]
    g1 = np.empty((b, n, 3), np.float32)
    g2 = np.empty((b, m, 3), np.float32)
    lib().orc_matchcostgrad(b, n, m, _pf(a), _pf(c), _pf(mt), _pf(g1), _pf(g2))
    return g1, g2


def knn(x, k: int, return_dist: bool = False):
    """x (B,C,N) channels-first -> idx (B,N,k) int64 ascending by (distance, index), self included."""
    a = _f(x)
    b, c, n = a.shape
    idx = np.empty((b, n, k), np.int64)
    dist = np.empty((b, n, k), np.float32) if return_dist else None
    lib().orc_knn(b, c, n, k, _pf(a), idx.ctypes.data_as(_L), _pf(dist) if return_dist else None)
    return (idx, dist) if return_dist else idx


def square_distance(t1, t2):
    """(B,N,C),(B,M,C) -> (B,N,M) canonical direct-form squared distances."""
    a, c = _f(t1), _f(t2)
    b, n, ch = a.shape
    m = c.shape[1]
    out = np.empty((b, n, m), np.float32)
    lib().orc_square_distance(b, n, m, ch, _pf(a), _pf(c), _pf(out))
    return out


def auction_emd(xyz1, xyz2, eps: float, iters: int):
    """-> dist (B,n) squared distance to the assigned target, assignment (B,n) i32, price (B,n)."""
    a, c = _f(xyz1), _f(xyz2)
    b, n, _ = a.shape
    m = c.shape[1]
    dist = np.zeros((b, n), np.float32)
    asg = np.full((b, n), -1, np.int32)
    price = np.zeros((b, n), np.float32)
    rc = lib().orc_auction_emd(b, n, m, _pf(a), _pf(c), float(eps), int(iters), _pf(dist), _pi(asg), _pf(price))
    if rc != 1:
        raise ValueError("auction_emd: shape rules violated (n==m, n%1024==0, B<=512)")
    return dist, asg, price


def auction_emd_grad(xyz1, xyz2, gdist, assignment):
    a, c, g, s = _f(xyz1), _f(xyz2), _f(gdist), _i(assignment)
    b, n, _ = a.shape
    g1 = np.empty((b, n, 3), np.float32)
    lib().orc_auction_emd_grad(b, n, _pf(a), _pf(c), _pf(g), _pi(s), _pf(g1))
    return g1
