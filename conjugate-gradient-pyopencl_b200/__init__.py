"""cg_b200 -- B200-native Conjugate Gradient engine behind the `cg()` entry point of
ziyamammadov/conjugate-gradient-pyopencl.

Layout
  csrc/        hand-written sm_100a CUDA kernels + the C ABI (include/clcg.h, include/cgb200.h)
  _lib.py      ctypes binding of liboclcg.so
  engine.py    `Matrix`: a resident CSR matrix with spmv() / solve()   (the handle API)
  cl.py        drop-in for the reference's `cl` module (cl.py:16-360): `import cl as pcl`
  problems.py  synthetic systems of the BASELINE.json configs
  sharded.py   row-block / RHS-split multi-GPU drivers

Import as `cg_b200` (see /cg_b200.py at the repo root; the directory name has a hyphen).
"""
from . import _lib, problems            # noqa: F401
from ._lib import CgError, F32, F64, C64, C128, LAYOUT_CLCG, LAYOUT_ROWMAJOR  # noqa: F401
from .engine import Matrix, PCG, cg, device_count  # noqa: F401

__all__ = ["Matrix", "PCG", "cg", "device_count", "problems", "CgError"]
