"""Is the 3e-9 distance between the device and the oracle after 12 CG iterations on the small power-law system the
system's own sensitivity to the summation order, or a kernel?  Every SpMV schedule against the oracle, numpy, and each other."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cg_b200, cg_b200.problems as P, cpu_ref, np_cg
A = P.powerlaw_spd(n=9000, nnz_target=120000, max_row=2500)
b = A @ np.linspace(-1, 1, A.shape[0])
rel = lambda a, c: float(np.linalg.norm(a - c) / np.linalg.norm(c))
for its in (4, 8, 12, 20):
    ref, _, _ = cpu_ref.cg(A.data, A.indptr, A.indices, b, iters=its)
    npx = np_cg.cg(A, b, x=np.zeros(A.shape[0]), maxit=its)
    xs = {}
    with cg_b200.Matrix.from_scipy(A) as M:
        M.set_option("solver", 1)
        for v in (1, 2, 3, 6):
            M.set_option("spmv_variant", v)
            xs[v] = M.solve(b, max_iterations=its)[0].copy()
        y = {v: None for v in (1, 2, 3, 6)}
        xr = np.linspace(-1, 1, A.shape[0])
        for v in (1, 2, 3, 6):
            M.set_option("spmv_variant", v)
            y[v] = M.spmv(xr).copy()
    print(its, "oracle vs numpy", rel(ref, npx), {v: (rel(x, ref), rel(x, npx), rel(x, xs[1])) for v, x in xs.items()},
          "spmv vs exact", {v: rel(yy, A @ xr) for v, yy in y.items()})
