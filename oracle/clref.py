"""oracle/clref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes loader for oracle/_ref/libclref.so: the reference's own OpenCL kernels (kernel/real/*.cl,
kernel/complex/*.cl, cmplx.h -- #included from /root/reference at build time) executed on the CPU through
oracle/clref/clref.cpp, inside the launch sequence of clcg.c.  Exists only where /root/reference does (the build
container); tests that use it are marked `reference`, fixtures made from it are committed under tests/golden/.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
_LIB = None


def available():
    return os.path.isdir(os.path.join(REFERENCE, "kernel", "real"))


def build(force=False):
    so = os.path.join(_HERE, "_ref", "libclref.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        vp, i = ctypes.c_void_p, ctypes.c_int
        L.clref_cg.argtypes = [i, i, vp, vp, vp, vp, vp, i, i, i]
        L.clref_cg.restype = i
        _LIB = L
    return _LIB


def cg(vals, rowptr, cols, b, x0=None, k=1, iters=10):
    """cg() of clcg.h:3-5 with the reference's kernels: float32 or complex64, n >= 256, k <= 4."""
    vals = np.ascontiguousarray(vals)
    dt = vals.dtype
    assert dt in (np.float32, np.complex64), "the reference has single precision only (main.c:49)"
    n = rowptr.size - 1
    b = np.ascontiguousarray(b, dtype=dt)
    x = np.zeros(n * k, dtype=dt) if x0 is None else np.array(x0, dtype=dt, copy=True, order="C")
    rowptr = np.ascontiguousarray(rowptr, dtype=np.intc)
    cols = np.ascontiguousarray(cols, dtype=np.intc)
    rc = lib().clref_cg(n, vals.size, vals.ctypes.data, b.ctypes.data, rowptr.ctypes.data, cols.ctypes.data,
                        x.ctypes.data, k, iters, int(dt.kind == "c"))
    if rc != 0:
        raise ValueError(f"clref_cg: {rc} (-1: k not in 1..4, -2: n < 256)")
    return x
