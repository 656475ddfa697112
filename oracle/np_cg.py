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
