"""ctypes binding of liboclcg.so: one Python function per symbol of include/clcg.h and include/cgb200.h.

There is no fallback: if the library cannot be built or loaded, importing fails loudly.
"""
import ctypes
import os

import numpy as np

try:
    from . import build as _build
except ImportError:  # imported as a top-level module (package dir on sys.path, like the reference's `import cl`)
    import build as _build

F32, F64, C64, C128 = 0, 1, 2, 3
LAYOUT_CLCG, LAYOUT_ROWMAJOR = 0, 1
FLAG_MAXIT, FLAG_BREAKDOWN = 1, 2
P2P_BLOB_BYTES = 256       # CGB200_P2P_BLOB_BYTES
DTYPE_CODE = {np.dtype(np.float32): F32, np.dtype(np.float64): F64,
              np.dtype(np.complex64): C64, np.dtype(np.complex128): C128}
CODE_DTYPE = {v: k for k, v in DTYPE_CODE.items()}

_vp, _i, _ll, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_double

# symbol -> (restype, argtypes); the complete exported surface
SIGNATURES = {
    "cg": (_vp, [_i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i]),
    "cgd": (_vp, [_i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i]),
    "cgb200_create": (_i, [ctypes.POINTER(_vp), _i, _ll, _vp, _vp, _vp, _i, _i]),
    "cgb200_create_grid": (_i, [ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "cgb200_read_matrix": (_i, [_vp, _vp, _vp, _vp]),
    "cgb200_update": (_i, [_vp, _vp, _vp, _vp]),
    "cgb200_destroy": (_i, [_vp]),
    "cgb200_set_stream": (_i, [_vp, _vp]),
    "cgb200_set_option": (_i, [_vp, ctypes.c_char_p, _ll]),
    "cgb200_get_option": (_i, [_vp, ctypes.c_char_p, ctypes.POINTER(_ll)]),
    "cgb200_spmv": (_i, [_vp, _vp, _vp, _i, _i]),
    "cgb200_solve": (_i, [_vp, _vp, _vp, _i, _i, _d, _vp, _vp, _vp, _i]),
    "cgb200_solve_pcg": (_i, [_vp, _vp, _vp, _vp, _i, _i, _d, _vp, _vp, _vp]),
    "cgb200_check_guards": (_ll, [_vp]),
    "cgb200_plan_march_runs": (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _i, ctypes.POINTER(_i)]),
    "cgb200_time_kernel": (_i, [_vp, _i, _i, _i, ctypes.POINTER(_d)]),
    "cgb200_last_timing": (_i, [_vp, ctypes.POINTER(_d)]),
    "cgb200_info": (_i, [_vp, ctypes.POINTER(_ll)]),
    "cgb200_read_trace": (_i, [_vp, _vp, _i]),
    "cgb200_debug_read_patterns": (_i, [_vp, _i, _vp, ctypes.c_size_t]),
    "cgb200_cg": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i]),
    "cgb200_clear_cache": (_i, []),
    "cgb200_nccl_unique_id": (_i, [_vp]),
    "cgb200_shard_create": (_i, [ctypes.POINTER(_vp), _i, _i, _vp, _i, _i, _i, _ll, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "cgb200_shard_destroy": (_i, [_vp]),
    "cgb200_shard_local": (_vp, [_vp]),
    "cgb200_shard_p2p_export": (_i, [_vp, _vp]),
    "cgb200_shard_p2p_import": (_i, [_vp, _vp, _vp]),
    "cgb200_shard_p2p_enable": (_i, [_vp, _i]),
    "cgb200_shard_set_stream": (_i, [_vp, _vp]),
    "cgb200_shard_set_option": (_i, [_vp, ctypes.c_char_p, _ll]),
    "cgb200_shard_solve": (_i, [_vp, _vp, _vp, _i, _d, ctypes.POINTER(_i), ctypes.POINTER(_d)]),
    "cgb200_shard_info": (_i, [_vp, ctypes.POINTER(_ll)]),
    "cgb200_shard_last_timing": (_i, [_vp, ctypes.POINTER(_d)]),
    "cgb200_last_error": (ctypes.c_char_p, []),
    "cgb200_device_count": (_i, []),
    "cgb200_version": (ctypes.c_char_p, []),
}

_LIB = None


class CgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cgb200 error {code}: {msg}")
        self.code = code


def lib():
    """Build (if stale) and load liboclcg.so."""
    global _LIB
    if _LIB is None:
        path = os.environ.get("CGB200_LIB") or _build.build()
        L = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the library lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        L._path = path
        _LIB = L
    return _LIB


def check(rc):
    if rc < 0:
        raise CgError(rc, lib().cgb200_last_error().decode(errors="replace"))
    return rc


def ptr(a):
    """void* of a numpy array, a torch tensor (host or device), an int address, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data_as(ctypes.c_void_p)
    if hasattr(a, "data_ptr"):
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return ctypes.c_void_p(a.data_ptr())
    raise TypeError(f"cannot take the address of {type(a)}")
