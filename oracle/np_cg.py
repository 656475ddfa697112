"""oracle/np_cg.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Double-precision oracle: the numpy CG of the reference
(`CG`, /root/reference/helmFE_var.py:507-544) restated so it can travel to the
GPU box (where /root/reference does not exist).  Same recurrence, same numpy /
scipy operations in the same order -- tests/test_oracle.py checks it is
bit-identical to the imported reference function at a fixed iteration count
(fixtures made by oracle/make_golden.py).

Like the reference: unconjugated `dot` (COCG for complex-symmetric A), no
breakdown guard, `tol` ignored unless `stop=True` (an instrumented variant used
for the iterations-to-convergence parity: stop at the first iteration with
sqrt(|delta_k| / |delta_0|) < tol, the quantity clcg.c:384-391 already forms).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this.
"""
import numpy as np


def cg(A, b, x=None, tol=1e-5, maxit=1000, stop=False, history=None):
    """Returns x (stop=False, as helmFE_var.CG) or (x, iterations) (stop=True)."""
    if x is None:
        x = np.zeros(b.size, dtype=complex)          # helmFE_var.py:508-509
    r = b - A.dot(x)                                 # :510-512
    d = r                                            # :514
    delta_new = np.dot(r, r)                         # :516
    delta_0 = abs(delta_new)
    if history is not None:
        history.append(delta_new)
    it = 0
    for it in range(1, maxit + 1):                   # :519
        q = A.dot(d)                                 # :520
        alpha = delta_new / np.dot(d, q)             # :522-524
        x = x + alpha * d                            # :527
        r = r - alpha * q                            # :530
        delta_old = delta_new                        # :533
        delta_new = np.dot(r, r)                     # :535
        if history is not None:
            history.append(delta_new)
        if stop and np.sqrt(abs(delta_new) / delta_0) < tol:
            break
        beta = delta_new / delta_old                 # :539
        d = r + beta * d                             # :542
    return (x, it) if stop else x


def cg_abs_tol(A, b, tol=1e-5, x=None, max_it=None):
    """The OTHER numpy CG of the reference: the drivers' local `CG` (/root/reference/p_h-PY_C-CL.py:1338-1369,
    used by as_prec when UseCG == 5, :1916-1923).  Same recurrence as `cg` above, but it ignores `maxit`
    (the loop bound is 2*b.size, :1349) and stops on the ABSOLUTE residual sqrt(|r.r|) < tol (:1364-1367).
    Same numpy operations in the same order, so results are bit-identical to the driver's
    (oracle/run_reference_driver.py asserts that when it makes tests/golden/asprec_*.npz).
    Returns (x, iterations performed)."""
    if x is None:
        x = np.zeros(b.size, dtype=complex)                  # :1346-1347
    r = b - A.dot(x)                                         # :1349
    d = None
    rho_prev = None
    it = 0
    for i in range(2 * b.size if max_it is None else max_it):   # :1350
        rho = np.dot(r, r)                                   # :1352-1353 (z = r, no preconditioner)
        d = r if i == 0 else r + (rho / rho_prev) * d        # :1355-1359
        q = A.dot(d)                                         # :1360
        alpha = rho / np.dot(d, q)                           # :1361
        x = x + alpha * d                                    # :1362
        r = r - alpha * q                                    # :1363
        it = i + 1
        if np.sqrt(abs(np.dot(r, r))) < tol:                 # :1364-1367
            break
        rho_prev = rho                                       # :1368
    return x, it


def pcg(A, b, M=None, x=None, tol=1e-6, maxit=1000):
    """The reference's preconditioned CG (/root/reference/helmFE_var.py:546-586), restated for the day the engine
    gets a preconditioner.  M: None (z = r, :557-558), a 1-D array of the INVERSE diagonal
    -- what the reference applies as a sparse matrix with one entry per row, z = M.dot(r), :559-563 -- or a
    callable (:566-567).  Unconjugated dots; stops when sqrt(|r.r|) < tol (:579-583).  Returns (x, i) like the
    reference: i is the index of the last iteration performed."""
    if M is not None and not callable(M):
        import scipy.sparse as sp
        M = sp.csr_matrix(sp.diags(np.asarray(M)))           # the same sparse product the reference performs
    if x is None:
        x = np.zeros(b.size, dtype=complex)                  # :553-554
    r = b - A.dot(x)                                         # :556
    p = None
    rho_prev = None
    i = 0
    for i in range(maxit):                                   # :557
        if M is None:
            z = r
        elif callable(M):
            z = M(r)
        else:
            z = M.dot(r)                                     # :563 (nnz <= n: "z = M.dot(r)")
        rho = np.dot(r, z)                                   # :568
        p = z if i == 0 else z + (rho / rho_prev) * p        # :570-574
        q = A.dot(p)                                         # :575
        alpha = rho / np.dot(p, q)                           # :576
        x = x + alpha * p                                    # :577
        r = r - alpha * q                                    # :578
        if np.sqrt(abs(np.dot(r, r))) < tol:                 # :579-583
            break
        rho_prev = rho                                       # :584
    return x, i
