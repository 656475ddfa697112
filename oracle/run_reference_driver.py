"""oracle/run_reference_driver.py -- TEST INFRASTRUCTURE.  Runs the reference's own Helmholtz driver,
UNMODIFIED, in the build container and records what it hands to the hot path and what comes back.

    python oracle/run_reference_driver.py [M_s W_s CGMaxIT] [--out DIR]       (default 2 12 40, tests/golden)

The driver (`/root/reference/p_h-PY_C-CL.py`) is executed as `__main__` with the same argv a user would
give it (p_h-PY_C-CL.py:2-17).  It needs mpi4py, pyopencl, an OpenCL device and ./build/liboclcg.so at
import; none of the first three exists in this image, so this script supplies
  * a one-rank `mpi4py.MPI` (COMM_WORLD of size 1; Isend/Irecv to self are a mailbox keyed by tag),
  * an empty `pyopencl` (the driver only enumerates platforms, p_h-PY_C-CL.py:88-93),
  * `cl` = this repo's drop-in module (conjugate-gradient-pyopencl_b200/cl.py) with the device call behind
    it (`cl._solve`) answered by the CPU oracle (oracle/cpu_ref.c) -- there is no GPU here --, and
  * `ctypes.CDLL("./build/liboclcg.so")` answered by an object whose `cg` is the same oracle and whose
    `connect()` is libc's, as with the reference library.
As shipped the driver runs its variants 0 (exact subdomain solves), 1 and 2 (pyopencl CG, single / multi
RHS) and 5 (its own numpy CG) one after the other (p_h-PY_C-CL.py:3622).  A profile hook records

  every call of the driver's numpy `CG(A, b, tol, maxit)` (p_h-PY_C-CL.py:1338-1369) with its result --
      the REFERENCE's own arithmetic on the reference's own subdomain systems --, and
  every call that reaches `pcl.CG(...)` (the arguments of the hot path exactly as as_prec builds them,
      p_h-PY_C-CL.py:1924-1937 and :1956-1969).

Two more modes run the other two drivers the same way (no fixtures, they print a summary that tests check):
    --multi-gpu        p_h-PY_C-CL-multi-GPU.py on two pretend devices; its own thread-per-device RHS split is then
                       driven with the recorded arrays (run_multi_gpu_script)
    --old-api <UseCG>  p_helmholtz.py: `pcl.create_kernels(1)`, 9-argument `pcl.CG`, `CDLL("./liboclcg.so")`

tests/golden/asprec_<M_s>_<W_s>.npz then holds, for the first preconditioner application:
  data/indices/indptr   the subdomain matrix P[0] (complex128, from the driver's local_rect)
  z                     [n_my][size] the right-hand sides z[p] (complex128)
  x_numpy_cg            [n_my][size] what the driver's numpy CG returned for them (tol 1e-5 on |r|)
  numpy_cg_iters        iterations that took (re-derived by the oracle restatement, checked to agree)
  cl_args_*             the arrays as_prec passed to pcl.CG in variant 2 (csingle / intc), n_rhs, n_iterations
  cl_result_reference_kernels   x of that very call computed by the reference's own OpenCL kernels (oracle/clref)
  gmres_iterations      outer iterations per variant, as printed by the driver
"""
import ctypes
import io
import os
import re
import runpy
import sys
import types
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
DRIVER = os.path.join(REF, "p_h-PY_C-CL.py")
sys.path.insert(0, HERE)


# --------------------------------------------------------------------------------------------
# one-rank mpi4py
# --------------------------------------------------------------------------------------------
class _Request:
    def __init__(self, comm=None, buf=None, tag=None):
        self.comm, self.buf, self.tag = comm, buf, tag

    def _complete(self):
        if self.buf is not None:
            msg = self.comm.mailbox.pop(self.tag)
            flat = self.buf.reshape(-1)
            flat[:msg.size] = msg
            self.buf = None

    @staticmethod
    def Wait(req):
        req._complete()

    @staticmethod
    def Waitall(reqs):
        for r in reqs:
            r._complete()


class _Comm:
    def __init__(self):
        self.mailbox = {}

    def Get_size(self):
        return 1

    def Get_rank(self):
        return 0

    def allreduce(self, v):
        return v

    def Barrier(self):
        pass

    def Isend(self, spec, dest=0, tag=0):
        self.mailbox[tag] = np.array(spec[0], copy=True).reshape(-1)
        return _Request()

    def Irecv(self, spec, source=0, tag=0):
        return _Request(self, spec[0], tag)


def install_stubs(record):
    import cpu_ref
    cpu_ref.build()

    mpi4py = types.ModuleType("mpi4py")
    MPI = types.ModuleType("mpi4py.MPI")
    MPI.COMM_WORLD = _Comm()
    MPI.COMPLEX = "COMPLEX"
    MPI.Request = _Request
    mpi4py.MPI = MPI
    sys.modules["mpi4py"] = mpi4py
    sys.modules["mpi4py.MPI"] = MPI

    pyopencl = types.ModuleType("pyopencl")
    pyopencl.get_platforms = lambda: []
    pyopencl.device_type = types.SimpleNamespace(GPU=4)
    sys.modules["pyopencl"] = pyopencl

    # the drop-in `cl` module, device call answered by the oracle
    sys.path.insert(0, os.path.join(ROOT, "conjugate-gradient-pyopencl_b200"))
    sys.modules.pop("cl", None)
    import cl as pcl

    def oracle_solve(dev, size, nnz, a_values, b_values, a_pointers, a_cols, x, n_rhs, n_iterations):
        record["cl_calls"].append(dict(size=size, nnz=nnz, a_values=a_values.copy(), b_values=b_values.copy(),
                                       a_pointers=a_pointers.copy(), a_cols=a_cols.copy(), x_in=x.copy(),
                                       n_rhs=n_rhs, n_iterations=n_iterations))
        out, _, _ = cpu_ref.cg(a_values, a_pointers, a_cols, b_values, x0=x, k=n_rhs, iters=n_iterations)
        x[...] = out
        return x

    pcl._solve = oracle_solve
    pcl.get_gpu_devices = lambda: [pcl.Device(0)]          # no CUDA device in the build container

    real_cdll = ctypes.CDLL

    class FakeLib:
        """CDLL("./build/liboclcg.so"): `cg` by the oracle, anything else (connect) from libc."""
        def __init__(self, *a, **k):
            self._libc = real_cdll(None)

            def cg(size, nnz, a_values, b_values, row_ptr, col_idx, x, n_rhs, n_it, is_complex):
                record["c_calls"] += 1
                out, _, _ = cpu_ref.cg(a_values, row_ptr, col_idx, b_values, x0=x, k=n_rhs, iters=n_it)
                x[...] = out
                return 0
            self.cg = _Settable(cg)

        def __getattr__(self, name):
            return getattr(self._libc, name)

    class _Settable:
        def __init__(self, fn):
            self.fn, self.argtypes, self.restype = fn, None, None

        def __call__(self, *a):
            return self.fn(*a)

    # only the drivers' own literals (p_h-PY_C-CL.py:38, p_helmholtz.py:29) get the stand-in
    ctypes.CDLL = lambda name=None, *a, **k: FakeLib() if name in ("./build/liboclcg.so", "./liboclcg.so") else real_cdll(name, *a, **k)


def profile_hook(record):
    """Records the driver's numpy CG calls: arguments on entry, the returned x on exit."""
    def hook(frame, event, arg):
        code = frame.f_code
        if code.co_name != "CG" or not code.co_filename.endswith("p_h-PY_C-CL.py"):
            return
        if event == "call":
            loc = frame.f_locals
            record["_open"][id(frame)] = dict(A=loc["A"], b=np.array(loc["b"], copy=True), tol=loc["tol"], maxit=loc["maxit"])
        elif event == "return" and id(frame) in record["_open"]:
            ent = record["_open"].pop(id(frame))
            ent["x"] = np.array(arg, copy=True)
            record["numpy_cg_calls"].append(ent)
    return hook


def run(M_s, W_s, maxit):
    record = {"cl_calls": [], "c_calls": 0, "numpy_cg_calls": [], "_open": {}}
    install_stubs(record)
    argv, cwd = sys.argv, os.getcwd()
    tmp = os.path.join("/tmp", f"refdrv_{os.getpid()}")
    os.makedirs(tmp, exist_ok=True)            # the driver writes output_*.txt into the CWD
    os.chdir(tmp)
    sys.argv = [DRIVER, str(M_s), str(W_s), "2", str(maxit)]
    buf = io.StringIO()
    sys.setprofile(profile_hook(record))
    try:
        with redirect_stdout(buf):
            runpy.run_path(DRIVER, run_name="__main__")
    except SystemExit:
        pass
    finally:
        sys.setprofile(None)
        sys.argv = argv
        os.chdir(cwd)
    record["stdout"] = buf.getvalue()
    return record


def run_multi_gpu_script(M_s=3, W_s=12, maxit=20, n_devices=2):
    """The reference's multi-GPU driver (p_h-PY_C-CL-multi-GPU.py), unmodified, on `n_devices` pretend devices.

    As shipped it only runs variant 2 (`cgs = [2]`, :3684); its RHS-split code -- distribute_workloads_on_devices
    (:2123-2140), run at import, and distribute_computations_with_threads (:2142-2181), variant 6 -- is then driven
    here with the very arrays as_prec built for variant 2: one thread per device calling this repo's
    cl.conjugate_gradient_multi_gpu on its contiguous block of right-hand sides.  Columns are independent CGs, so
    the stitched result must equal the single multi-RHS call bit for bit.  Returns a summary dict."""
    record = {"cl_calls": [], "c_calls": 0, "numpy_cg_calls": [], "_open": {}}
    install_stubs(record)
    import cl as pcl
    pcl.get_gpu_devices = lambda: [pcl.Device(i) for i in range(n_devices)]
    script = os.path.join(REF, "p_h-PY_C-CL-multi-GPU.py")
    argv, cwd = sys.argv, os.getcwd()
    tmp = os.path.join("/tmp", f"refdrv_{os.getpid()}")
    os.makedirs(tmp, exist_ok=True)
    os.chdir(tmp)
    sys.argv = [script, str(M_s), str(W_s), "2", str(maxit)]
    try:
        with redirect_stdout(io.StringIO()) as buf:
            ns = runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv = argv
        os.chdir(cwd)
    n_my = M_s * M_s
    workloads = ns["workloads"]
    ranges = sorted((w[0], w[1]) for w in workloads.values())
    multi = [c for c in record["cl_calls"] if c["n_rhs"] == n_my]
    assert multi, "variant 2 did not reach pcl.CG"
    c2 = multi[0]
    import cpu_ref
    whole, _, _ = cpu_ref.cg(c2["a_values"], c2["a_pointers"], c2["a_cols"], c2["b_values"], x0=c2["x_in"], k=n_my,
                             iters=c2["n_iterations"])
    before = len(record["cl_calls"])
    x = c2["x_in"].copy()
    ns["distribute_computations_with_threads"](c2["size"], c2["nnz"], c2["a_values"], c2["b_values"], c2["a_pointers"],
                                               c2["a_cols"], x, n_my, c2["n_iterations"], workloads)
    split_calls = record["cl_calls"][before:]
    return {"ranges": ranges, "n_my": n_my, "devices": sorted(d.ordinal for d in workloads),
            "split_calls": sorted((c["n_rhs"]) for c in split_calls), "identical": bool(np.array_equal(x, whole)),
            "gmres_iterations": [int(m) for m in re.findall(r"####it:\s*(\d+)", buf.getvalue())]}


def run_old_api_script(use_cg, M_s=2, W_s=12):
    """p_helmholtz.py, the oldest of the three drivers, unmodified: `pcl.create_kernels(1)` at import (:31) and the
    9-argument `pcl.CG(size, nnz, a, b, ptr, cols, x, n_rhs, maxit)` (:1839, :1873); its numpy CG is imported from
    helmFE_var (:25).  Returns how often and with how many right-hand sides the hot path was called."""
    record = {"cl_calls": [], "c_calls": 0, "numpy_cg_calls": [], "_open": {}}
    install_stubs(record)
    sys.path.insert(0, REF)                    # `from helmFE_var import CG`
    script = os.path.join(REF, "p_helmholtz.py")
    argv, cwd = sys.argv, os.getcwd()
    tmp = os.path.join("/tmp", f"refdrv_{os.getpid()}")
    os.makedirs(tmp, exist_ok=True)
    os.chdir(tmp)
    sys.argv = [script, str(M_s), str(W_s), str(use_cg)]
    buf = io.StringIO()
    try:
        with redirect_stdout(buf):
            runpy.run_path(script, run_name="__main__")
    except SystemExit:
        pass
    finally:
        sys.argv = argv
        os.chdir(cwd)
    return {"use_cg": use_cg, "cl_calls": len(record["cl_calls"]), "n_rhs": sorted({c["n_rhs"] for c in record["cl_calls"]}),
            "c_calls": record["c_calls"], "gmres_iterations": [int(m) for m in re.findall(r"####it:\s*(\d+)", buf.getvalue())]}


def main():
    if "--old-api" in sys.argv:
        print(run_old_api_script(int(sys.argv[sys.argv.index("--old-api") + 1])))
        return
    if "--multi-gpu" in sys.argv:
        print(run_multi_gpu_script())
        return
    out_dir = os.path.join(ROOT, "tests", "golden")
    if "--out" in sys.argv:
        i = sys.argv.index("--out")
        out_dir = sys.argv[i + 1]
        del sys.argv[i:i + 2]
    M_s, W_s, maxit = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (2, 12, 40)
    rec = run(M_s, W_s, maxit)
    out = rec["stdout"]
    print(out[-1500:])
    n_my = M_s * M_s
    assert len(rec["numpy_cg_calls"]) >= n_my, "variant 5 (numpy CG) did not run"
    multi = [c for c in rec["cl_calls"] if c["n_rhs"] == n_my]
    assert multi, "variant 2 (multi-RHS pcl.CG) did not run"
    first = rec["numpy_cg_calls"][:n_my]
    A = first[0]["A"].tocsr()
    A.sort_indices()
    import np_cg
    iters = []
    for c in first:          # the restatement, with the driver-CG's stopping rule, must reproduce the reference bit for bit
        x, it = np_cg.cg_abs_tol(A, c["b"], tol=c["tol"])
        assert np.array_equal(x, c["x"]), "oracle/np_cg.py differs from the driver's CG"
        iters.append(it)
    c2 = multi[0]
    # the same call -- the very arrays as_prec built -- through the reference's OWN OpenCL kernels (oracle/clref);
    # the columns are independent CGs, so they go through in groups of <= 4 (the instantiated N_RHS values)
    import clref
    size = c2["size"]
    x_kernels = np.zeros_like(c2["x_in"])
    for c0 in range(0, n_my, 4):
        kk = min(4, n_my - c0)
        x_kernels[c0 * size:(c0 + kk) * size] = clref.cg(c2["a_values"], c2["a_pointers"], c2["a_cols"],
                                                         c2["b_values"][c0 * size:(c0 + kk) * size],
                                                         x0=c2["x_in"][c0 * size:(c0 + kk) * size], k=kk,
                                                         iters=c2["n_iterations"])
    gm = [int(m) for m in re.findall(r"####it:\s*(\d+)", out)]          # one per variant 0, 1, 2, 5 (p_h-PY_C-CL.py:3622)
    dst = os.path.join(out_dir, f"asprec_{M_s}_{W_s}.npz")
    np.savez_compressed(
        dst, data=A.data, indices=A.indices, indptr=A.indptr, n=A.shape[0],
        z=np.stack([c["b"] for c in first]), x_numpy_cg=np.stack([c["x"] for c in first]),
        numpy_cg_tol=first[0]["tol"], numpy_cg_iters=np.array(iters),
        cl_args_a_values=c2["a_values"], cl_args_b_values=c2["b_values"], cl_args_a_pointers=c2["a_pointers"],
        cl_args_a_cols=c2["a_cols"], cl_args_x_in=c2["x_in"], cl_args_n_rhs=c2["n_rhs"],
        cl_args_n_iterations=c2["n_iterations"], cl_args_size=c2["size"], cl_args_nnz=c2["nnz"],
        cl_result_reference_kernels=x_kernels,
        cl_calls_total=len(rec["cl_calls"]), numpy_cg_calls_total=len(rec["numpy_cg_calls"]),
        gmres_iterations=np.array(gm))
    print("wrote", dst, "n =", A.shape[0], "nnz =", A.nnz, "n_my =", n_my, "numpy CG iterations", iters,
          "pcl.CG calls", len(rec["cl_calls"]))


if __name__ == "__main__":
    main()
