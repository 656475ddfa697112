"""Host-side mirror of the C ABI: a resident CSR matrix and the calls on it.

`cg(...)` has the argument list of the reference's entry point (clcg.h:3-5) and goes
through the exported `cg` / `cgd` symbols, exactly as the drivers' ctypes call does
(p_h-PY_C-CL.py:1948-1950).  `Matrix` wraps a `cgb200_handle`.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import check, lib, ptr


def device_count():
    return lib().cgb200_device_count()


def cg(size, non_zeros, a_values, b, a_pointers, a_cols, x, n_rhs, n_iterations, is_complex=None):
    """`cg()` / `cgd()` of include/clcg.h on numpy arrays; x is updated in place and returned.

    float32/complex64 arrays go to `cg` (the reference's precision), float64/complex128 to `cgd`."""
    a_values = np.ascontiguousarray(a_values)
    dt = a_values.dtype
    if dt not in _lib.DTYPE_CODE:
        raise TypeError(f"unsupported value dtype {dt}")
    cplx = dt.kind == "c"
    if is_complex is not None and bool(is_complex) != cplx:
        raise ValueError("is_complex does not match the dtype of a_values")
    b = np.ascontiguousarray(b, dtype=dt)
    if x.dtype != dt or not x.flags["C_CONTIGUOUS"]:
        raise TypeError("x must be a C-contiguous array of the matrix dtype (it is written in place)")
    a_pointers = np.ascontiguousarray(a_pointers, dtype=np.intc)
    a_cols = np.ascontiguousarray(a_cols, dtype=np.intc)
    fn = lib().cg if dt.itemsize // (2 if cplx else 1) == 4 else lib().cgd
    ret = fn(int(size), int(non_zeros), ptr(a_values), ptr(b), ptr(a_pointers), ptr(a_cols), ptr(x),
             int(n_rhs), int(n_iterations), 1 if cplx else 0)
    if not ret:
        raise _lib.CgError(-2, lib().cgb200_last_error().decode(errors="replace"))
    return x


class SolveInfo:
    __slots__ = ("flags", "iterations", "relres", "delta_hist", "timing_ms")

    def __repr__(self):
        return (f"SolveInfo(flags={self.flags}, iterations={self.iterations.tolist()}, "
                f"relres={self.relres.tolist()}, timing_ms={self.timing_ms})")


def read_trace(handle, iterations):
    out = np.zeros((int(iterations), 8), dtype=np.uint64)
    check(lib().cgb200_read_trace(handle, out.ctypes.data_as(ctypes.c_void_p), int(iterations)))
    return out


class Matrix:
    """A CSR matrix resident in the HBM of one B200 (`cgb200_create`)."""

    def __init__(self, values, rowptr, cols, n=None, device=0, dtype=None):
        if isinstance(values, np.ndarray):
            values = np.ascontiguousarray(values if dtype is None else values.astype(dtype, copy=False))
            self.dtype = values.dtype
            nnz = values.size
            rowptr = np.ascontiguousarray(rowptr, dtype=np.intc)
            cols = np.ascontiguousarray(cols, dtype=np.intc)
            n = rowptr.size - 1 if n is None else n
        else:  # torch tensors already on the device
            import torch
            tmap = {torch.float32: np.float32, torch.float64: np.float64,
                    torch.complex64: np.complex64, torch.complex128: np.complex128}
            self.dtype = np.dtype(tmap[values.dtype])
            nnz = values.numel()
            n = rowptr.numel() - 1 if n is None else n
        self.n, self.nnz, self.device = int(n), int(nnz), int(device)
        self.code = _lib.DTYPE_CODE[np.dtype(self.dtype)]
        h = ctypes.c_void_p()
        check(lib().cgb200_create(ctypes.byref(h), self.n, self.nnz, ptr(values), ptr(rowptr), ptr(cols),
                                  self.code, self.device))
        self._h = h

    @classmethod
    def from_scipy(cls, A, dtype=None, device=0):
        A = A.tocsr()
        vals = A.data if dtype is None else A.data.astype(dtype)
        return cls(vals, A.indptr, A.indices, n=A.shape[0], device=device)

    def update(self, values, rowptr, cols):
        """New content with the same n / nnz / dtype (`cgb200_update`): no reallocation."""
        values = np.ascontiguousarray(values, dtype=self.dtype)
        rowptr = np.ascontiguousarray(rowptr, dtype=np.intc)
        cols = np.ascontiguousarray(cols, dtype=np.intc)
        if values.size != self.nnz or rowptr.size != self.n + 1 or cols.size != self.nnz:
            raise ValueError("update() needs arrays of the sizes the matrix was created with")
        check(lib().cgb200_update(self._h, ptr(values), ptr(rowptr), ptr(cols)))

    def close(self):
        if getattr(self, "_h", None):
            lib().cgb200_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- options / facts --------------------------------------------------------
    def set_option(self, key, value):
        check(lib().cgb200_set_option(self._h, key.encode(), int(value)))

    def get_option(self, key):
        v = ctypes.c_longlong()
        check(lib().cgb200_get_option(self._h, key.encode(), ctypes.byref(v)))
        return v.value

    def set_stream(self, cuda_stream):
        check(lib().cgb200_set_stream(self._h, ctypes.c_void_p(int(cuda_stream) if cuda_stream else 0)))

    def info(self):
        out = (ctypes.c_longlong * 10)()
        check(lib().cgb200_info(self._h, out))
        keys = ("n", "nnz", "dtype", "lanes_per_row", "spmv_grid", "sm_count", "launches", "graph_launches",
                "max_row", "device")
        return dict(zip(keys, list(out)))

    TRACE_EVENTS = ("spmv_start", "spmv_all_done", "spmv_end", "xr_start", "xr_all_done", "xr_end", "d_start",
                    "halo_ready")

    def read_trace(self, iterations):
        """Timeline of the last solve (set_option("trace", n) first): uint64 [iterations][8] nanoseconds,
        columns TRACE_EVENTS, 0 = not recorded."""
        return read_trace(self._h, iterations)

    KERNELS = {"spmv_dot": 0, "update_xr": 1, "update_d": 2, "spmv": 3, "dir_spmv": 4, "update_r": 5}

    def time_kernel(self, which, k=1, reps=50):
        """Mean milliseconds of one kernel of the CG loop over `reps` back-to-back launches (CUDA events)."""
        ms = ctypes.c_double()
        check(lib().cgb200_time_kernel(self._h, self.KERNELS[which], int(k), int(reps), ctypes.byref(ms)))
        return ms.value

    # -- operations ---------------------------------------------------------------
    def spmv(self, x, y=None, k=1, layout=_lib.LAYOUT_CLCG):
        """y = A x.  numpy in -> numpy out (blocking); device tensors in -> asynchronous, y required."""
        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(x, dtype=self.dtype)
            if y is None:
                y = np.empty_like(x)
        elif y is None:
            raise ValueError("y is required for device pointers")
        check(lib().cgb200_spmv(self._h, ptr(x), ptr(y), int(k), int(layout)))
        return y

    def solve(self, b, x=None, k=1, max_iterations=1000, tol=0.0, history=False, layout=_lib.LAYOUT_CLCG):
        """CG on k right-hand sides.  Returns (x, SolveInfo); x (initial guess) is updated in place."""
        if isinstance(b, np.ndarray):
            b = np.ascontiguousarray(b, dtype=self.dtype)
            if x is None:
                x = np.zeros(self.n * k, dtype=self.dtype)
            elif x.dtype != self.dtype or not x.flags["C_CONTIGUOUS"]:
                raise TypeError("x must be a C-contiguous array of the matrix dtype")
            if b.size != self.n * k or x.size != self.n * k:
                raise ValueError("b and x must hold k blocks of n values")
        elif x is None:
            raise ValueError("x is required for device pointers")
        its = np.zeros(k, dtype=np.intc)
        rel = np.zeros(k, dtype=np.float64)
        ncomp = 2 if np.dtype(self.dtype).kind == "c" else 1
        hist = np.zeros((max_iterations + 1, k, ncomp)) if history else None
        rc = check(lib().cgb200_solve(self._h, ptr(b), ptr(x), int(k), int(max_iterations), float(tol),
                                      ptr(its), ptr(rel), ptr(hist), int(layout)))
        info = SolveInfo()
        info.flags, info.iterations, info.relres = rc, its, rel
        if hist is not None:
            hist = hist[..., 0] + 1j * hist[..., 1] if ncomp == 2 else hist[..., 0]
        info.delta_hist = hist
        ms = (ctypes.c_double * 4)()
        check(lib().cgb200_last_timing(self._h, ms))
        info.timing_ms = dict(zip(("h2d", "init", "iterations", "d2h"), list(ms)))
        return x, info

    def solve_pcg(self, b, x=None, k=1, M_inv_diag=None, max_iterations=1000, tol=0.0, history=False):
        """Jacobi-preconditioned CG (`cgb200_solve_pcg`; the reference's helmFE_var.PCG with M an inverse diagonal).
        M_inv_diag: n values, None = 1 / diag(A).  tol is ABSOLUTE on sqrt|r.r|, as the reference stops.
        Returns (x, SolveInfo) with relres = sqrt|r.r| at exit and delta_hist = the history of r.r."""
        if isinstance(b, np.ndarray):
            b = np.ascontiguousarray(b, dtype=self.dtype)
            if x is None:
                x = np.zeros(self.n * k, dtype=self.dtype)
            elif x.dtype != self.dtype or not x.flags["C_CONTIGUOUS"]:
                raise TypeError("x must be a C-contiguous array of the matrix dtype")
        elif x is None:
            raise ValueError("x is required for device pointers")
        if isinstance(M_inv_diag, np.ndarray):
            M_inv_diag = np.ascontiguousarray(M_inv_diag, dtype=self.dtype)
            if M_inv_diag.size != self.n:
                raise ValueError("M_inv_diag must hold n values")
        its = np.zeros(k, dtype=np.intc)
        res = np.zeros(k, dtype=np.float64)
        ncomp = 2 if np.dtype(self.dtype).kind == "c" else 1
        hist = np.zeros((max_iterations + 1, k, ncomp)) if history else None
        rc = check(lib().cgb200_solve_pcg(self._h, ptr(M_inv_diag), ptr(b), ptr(x), int(k), int(max_iterations), float(tol),
                                          ptr(its), ptr(res), ptr(hist)))
        info = SolveInfo()
        info.flags, info.iterations, info.relres = rc, its, res
        if hist is not None:
            hist = hist[..., 0] + 1j * hist[..., 1] if ncomp == 2 else hist[..., 0]
        info.delta_hist = hist
        ms = (ctypes.c_double * 4)()
        check(lib().cgb200_last_timing(self._h, ms))
        info.timing_ms = dict(zip(("h2d", "init", "iterations", "d2h"), list(ms)))
        return x, info


def PCG(A, b, M=None, x=None, tol=1e-6, maxit=1000, verbose=False, device=0):
    """The calling convention of the reference's `PCG(A, b, M, x, tol, maxit)` (helmFE_var.py:546-586) on the B200:
    A scipy sparse, M None (plain CG recurrence with the PCG stopping rule), a 1-D array of the inverse diagonal, or a
    scipy sparse matrix with one entry per row (what the reference multiplies with, :559-563).  Returns (x, i) with i
    the index of the last iteration performed, like the reference."""
    import scipy.sparse as sp
    A = sp.csr_matrix(A)
    A.sort_indices()
    dt = np.result_type(A.dtype, np.asarray(b).dtype, np.complex128 if x is None else np.asarray(x).dtype)   # x defaults to complex zeros (:553)
    dt = np.dtype(dt)
    if M is None:
        dinv = np.ones(A.shape[0], dtype=dt)
    elif sp.issparse(M):
        if M.nnz > M.shape[0]:
            raise NotImplementedError("only an inverse-diagonal M runs on the device (helmFE_var.py:560-561 would call spsolve)")
        dinv = np.asarray(sp.csr_matrix(M).diagonal(), dtype=dt)
    elif callable(M) or isinstance(M, float):
        raise NotImplementedError("callable / inner-CG preconditioners (helmFE_var.py:564-567) are host-side constructs")
    else:
        dinv = np.asarray(M, dtype=dt)
    x0 = np.zeros(A.shape[0], dtype=dt) if x is None else np.ascontiguousarray(x, dtype=dt).copy()
    with Matrix(A.data.astype(dt), A.indptr, A.indices, n=A.shape[0], device=device) as Mx:
        xs, info = Mx.solve_pcg(np.asarray(b, dtype=dt), x=x0, M_inv_diag=dinv, max_iterations=int(maxit), tol=float(tol))
    return xs, max(int(info.iterations[0]) - 1, 0)
