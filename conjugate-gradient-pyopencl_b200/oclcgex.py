"""`oclcgex` -- the reference's example executable (main.c:13-61) on the B200 engine.

    python oclcgex.py <input matrix file> <number of RHS> <is complex> <number of iterations>

Same four arguments (main.c:15-18).  What main.c does, step by step:
  load a Matrix Market file (BeBOP SMC there, scipy here)            main.c:20
  expand symmetric storage to full storage                           main.c:25
  convert to CSR                                                     main.c:27
  b[r*n + i] = 5 (r + 1), x0 = 0                                      main.c:41-46
  narrow the values to single precision                              main.c:49-53
  cg(n, nnz, aValues, b, rowptr, colidx, x, nRHS, nIterations, isComplex)    main.c:56
main.c allocates complex buffers whatever <is complex> says and is only correct for 1; here a real
matrix with <is complex> = 0 runs the real float path.  Unlike main.c the residual is printed.
Extra: --double solves in double precision through cgd().
"""
import sys

import numpy as np


def load_csr(path):
    import scipy.io
    import scipy.sparse as sp
    A = scipy.io.mmread(path)          # symmetric / hermitian files come back fully expanded (main.c:25)
    A = sp.csr_matrix(A)               # main.c:27
    A.sum_duplicates()
    A.sort_indices()
    return A


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    double = "--double" in argv
    if double:
        argv.remove("--double")
    if len(argv) != 4:
        sys.stderr.write("Usage: ./CG <input matrix file> <number of RHS> <is complex> <number of iterations>\n")
        return 1
    path, n_rhs, is_complex, n_iter = argv[0], int(argv[1]), int(argv[2]), int(argv[3])
    try:
        A = load_csr(path)
    except Exception as e:             # main.c:21-24
        print("Could not read matrix", e)
        return 1
    if np.iscomplexobj(A.data) and not is_complex:
        print("matrix is complex: pass <is complex> = 1")
        return 1
    n = A.shape[0]
    if is_complex:
        dt = np.complex128 if double else np.complex64
    else:
        dt = np.float64 if double else np.float32
    a_values = A.data.astype(dt)                                            # main.c:50-53
    b = np.concatenate([np.full(n, 5.0 * (r + 1), dtype=dt) for r in range(n_rhs)])   # main.c:41-46
    x = np.zeros(n * n_rhs, dtype=dt)
    try:
        from . import engine
    except ImportError:
        import engine
    engine.cg(n, A.nnz, a_values, b, A.indptr, A.indices, x, n_rhs, n_iter)  # main.c:56
    Aw = A.astype(np.complex128 if is_complex else np.float64)
    for r in range(n_rhs):
        xr = x[r * n:(r + 1) * n].astype(Aw.dtype)
        res = np.linalg.norm(Aw @ xr - b[r * n:(r + 1) * n]) / np.linalg.norm(b[r * n:(r + 1) * n])
        print(f"rhs {r}: relative residual {res:.3e} after {n_iter} iterations")
    return 0


if __name__ == "__main__":
    sys.exit(main())
