"""oracle/cpu_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes loader for oracle/libcpu_ref.so (the C restatement of clcg.c + kernels).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_CODES = {np.dtype(np.float32): 0, np.dtype(np.float64): 1,
          np.dtype(np.complex64): 2, np.dtype(np.complex128): 3}


def build(force=False):
    so = os.path.join(_HERE, "libcpu_ref.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        vp, ip = ctypes.c_void_p, ctypes.c_void_p
        L.cpu_ref_cg.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ip, ip, vp,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_double, ip, vp]
        L.cpu_ref_cg.restype = ctypes.c_int
        L.cpu_ref_spmv.argtypes = [ctypes.c_int, ctypes.c_int, vp, ip, ip, vp, vp, ctypes.c_int]
        L.cpu_ref_spmv.restype = ctypes.c_int
        L.cpu_ref_threads.restype = ctypes.c_int
        L.cpu_ref_set_threads.argtypes = [ctypes.c_int]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def cg(vals, rowptr, cols, b, x0=None, k=1, iters=100, tol=0.0, want_hist=False):
    """Run the oracle.  b, x0: flat arrays of k blocks of n values (RHS r at r*n, clcg ABI).

    Returns (x, iterations_per_rhs, delta_hist or None)."""
    vals = np.ascontiguousarray(vals)
    dt = vals.dtype
    n = rowptr.size - 1
    b = np.ascontiguousarray(b, dtype=dt)
    x = np.zeros(n * k, dtype=dt) if x0 is None else np.array(x0, dtype=dt, copy=True, order="C")
    assert b.size == n * k and x.size == n * k
    rowptr = np.ascontiguousarray(rowptr, dtype=np.intc)
    cols = np.ascontiguousarray(cols, dtype=np.intc)
    its = np.zeros(k, dtype=np.intc)
    ncomp = 2 if dt.kind == "c" else 1
    hist = np.zeros((iters + 1, k, ncomp)) if want_hist else None
    rc = lib().cpu_ref_cg(_CODES[dt], n, vals.size, _p(vals), _p(b), _p(rowptr), _p(cols), _p(x),
                          k, iters, float(tol), _p(its), _p(hist))
    if rc:
        raise RuntimeError(f"cpu_ref_cg failed: {rc}")
    if hist is not None and ncomp == 2:
        hist = hist[..., 0] + 1j * hist[..., 1]
    elif hist is not None:
        hist = hist[..., 0]
    return x, its, hist


def pcg(vals, rowptr, cols, b, dinv, x0=None, k=1, iters=100, want_hist=False):
    """Jacobi-preconditioned CG in the device's arrangement (cpu_ref_impl.h::cpu_ref_pcg): exactly `iters` iterations.
    dinv: the n inverse diagonal entries.  Returns (x, r.r history or None)."""
    vals = np.ascontiguousarray(vals)
    dt = vals.dtype
    n = rowptr.size - 1
    b = np.ascontiguousarray(b, dtype=dt)
    dinv = np.ascontiguousarray(dinv, dtype=dt)
    x = np.zeros(n * k, dtype=dt) if x0 is None else np.array(x0, dtype=dt, copy=True, order="C")
    rowptr = np.ascontiguousarray(rowptr, dtype=np.intc)
    cols = np.ascontiguousarray(cols, dtype=np.intc)
    ncomp = 2 if dt.kind == "c" else 1
    hist = np.zeros((iters + 1, k, ncomp)) if want_hist else None
    L = lib()
    L.cpu_ref_pcg.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 6 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.cpu_ref_pcg.restype = ctypes.c_int
    rc = L.cpu_ref_pcg(_CODES[dt], n, _p(vals), _p(b), _p(rowptr), _p(cols), _p(x), _p(dinv), k, iters, _p(hist))
    if rc:
        raise RuntimeError(f"cpu_ref_pcg failed: {rc}")
    if hist is not None:
        hist = hist[..., 0] + 1j * hist[..., 1] if ncomp == 2 else hist[..., 0]
    return x, hist


def spmv(vals, rowptr, cols, x, k=1):
    vals = np.ascontiguousarray(vals)
    dt = vals.dtype
    n = rowptr.size - 1
    x = np.ascontiguousarray(x, dtype=dt)
    y = np.empty(n * k, dtype=dt)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.intc)
    cols = np.ascontiguousarray(cols, dtype=np.intc)
    lib().cpu_ref_spmv(_CODES[dt], n, _p(vals), _p(rowptr), _p(cols), _p(x), _p(y), k)
    return y


def threads():
    return lib().cpu_ref_threads()
