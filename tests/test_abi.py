"""CPU-side checks of the drop-in boundary: the library builds, loads and exports exactly the
symbols include/*.h declare; no compute call is made without a GPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import cg_b200
from cg_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in ("clcg.h", "cgb200.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("#"))
        names |= set(re.findall(r"CGB200_API[^;(]*?\b(\w+)\s*\(", src))
    return names


def test_every_declared_symbol_is_exported_and_bound():
    declared = _declared()
    assert {"cg", "cgd", "cgb200_create", "cgb200_solve", "cgb200_spmv", "cgb200_destroy"} <= declared
    L = _lib.lib()
    out = subprocess.check_output(["nm", "-D", "--defined-only", L._path], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert declared <= exported, declared - exported
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    # the library must not define `connect`: the drivers call libcg.connect() and rely on it
    # resolving to libc, and NCCL's bootstrap uses the socket call
    assert "connect" not in exported
    # nothing but the C ABI leaks out
    assert all(s in declared for s in exported), exported - declared


def test_copies_where_the_reference_drivers_look():
    _lib.lib()
    assert os.path.exists(os.path.join(ROOT, "build", "liboclcg.so"))   # p_h-PY_C-CL.py:38
    assert os.path.exists(os.path.join(ROOT, "liboclcg.so"))            # p_helmholtz.py:29


def test_reference_ctypes_binding_shape():
    """The exact argtypes the drivers install (p_h-PY_C-CL.py:1948-1950) bind to our `cg`."""
    from numpy.ctypeslib import ndpointer
    L = ctypes.CDLL(os.path.join(ROOT, "build", "liboclcg.so"))
    L.cg.argtypes = [ctypes.c_int, ctypes.c_int, ndpointer(dtype=np.csingle, ndim=1, flags="C"),
                     ndpointer(dtype=np.csingle, ndim=1, flags="C"), ndpointer(dtype=np.intc, ndim=1, flags="C"),
                     ndpointer(dtype=np.intc, ndim=1, flags="C"), ndpointer(dtype=np.csingle, ndim=1, flags="C"),
                     ctypes.c_int, ctypes.c_int, ctypes.c_int]
    assert L.cg is not None
    # libcg.connect() of the drivers (p_h-PY_C-CL.py:39) must fall through to libc
    assert hasattr(L, "connect")


def test_argument_validation_without_gpu():
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.cgb200_create(ctypes.byref(h), 0, 0, None, None, None, 0, 0) == -1
    assert b"bad matrix" in L.cgb200_last_error()
    assert L.cgb200_create(None, 4, 4, None, None, None, 0, 0) == -1
    assert L.cgb200_solve(None, None, None, 1, 1, 0.0, None, None, None, 0) == -1
    assert L.cgb200_spmv(None, None, None, 1, 0) == -1
    assert L.cgb200_set_option(None, b"x", 1) == -1
    assert L.cgb200_destroy(None) == 0
    assert L.cg(0, 0, None, None, None, None, None, 1, 1, 0) is None
    assert L.cgb200_version().startswith(b"cgb200")


def test_no_cpu_fallback_when_no_device():
    if cg_b200.device_count() > 0:
        pytest.skip("a GPU is present")
    A = cg_b200.problems.poisson2d(4)
    with pytest.raises(cg_b200.CgError):
        cg_b200.Matrix.from_scipy(A)
    x = np.zeros(16)
    with pytest.raises(cg_b200.CgError):
        cg_b200.cg(16, A.nnz, A.data, np.ones(16), A.indptr, A.indices, x, 1, 3)


def test_product_does_not_touch_the_oracle():
    """The shipped path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "conjugate-gradient-pyopencl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "cpu_ref" not in src and "np_cg" not in src and "oracle/" not in src, f
    out = subprocess.check_output(["ldd", _lib.lib()._path], text=True)
    assert "cpu_ref" not in out


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype of include/*.h against cg_b200._lib.SIGNATURES: the number of parameters and, per parameter,
    pointer / 64-bit integer / int / double -- a ctypes table that drifted from the header corrupts the call silently."""
    protos = {}
    for h in ("clcg.h", "cgb200.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("#"))
        for ret, name, params in re.findall(r"CGB200_API\s+([^;(]*?)\b(\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
            params = " ".join(params.split())
            protos[name] = (ret.strip(), [] if params in ("", "void") else [p.strip() for p in params.split(",")])

    def kind(c_decl):
        if "*" in c_decl or "[" in c_decl or re.search(r"\bcgb200_(handle|shard)\b", c_decl):
            return "ptr"
        if "long long" in c_decl or "size_t" in c_decl:
            return "i64"
        if "double" in c_decl:
            return "f64"
        return "i32"

    def ckind(ct):
        if ct in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(ct, "contents") or isinstance(ct, type(ctypes.POINTER(ctypes.c_int))):
            return "ptr"
        if ct in (ctypes.c_longlong, ctypes.c_size_t, ctypes.c_ulonglong):
            return "i64"
        if ct is ctypes.c_double:
            return "f64"
        return "i32"

    assert set(protos) == set(_lib.SIGNATURES)
    for name, (ret, params) in protos.items():
        res, args = _lib.SIGNATURES[name]
        assert len(params) == len(args), (name, params, args)
        for p, a in zip(params, args):
            assert kind(p) == ckind(a), (name, p, a)
        assert kind(ret + " ") == ckind(res), (name, ret, res)


def test_every_option_key_is_documented_in_the_header():
    """cgb200_set_option keys (the strcmp chain of option_slot) against the option list in include/cgb200.h."""
    src = open(os.path.join(ROOT, "conjugate-gradient-pyopencl_b200", "csrc", "cgb200.cu")).read()
    keys = set(re.findall(r'strcmp\(key, "(\w+)"\)', src))
    header = open(os.path.join(ROOT, "include", "cgb200.h")).read()
    assert len(keys) >= 15
    assert not [k for k in keys if f'"{k}"' not in header]
