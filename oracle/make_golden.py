"""oracle/make_golden.py -- TEST INFRASTRUCTURE.  Regenerates tests/golden/*.

Runs ONLY in the build container, where the reference is mounted at
/root/reference: imports the reference's own numpy CG and matrix generators
(helmFE_var.py imports with numpy/scipy alone; `Poisson` is lifted out of
p_helmholtz.py by `ast`, because that module needs mpi4py and a CDLL at import)
and stores their inputs and outputs as small fixtures, so that the tests that
pin the oracles can run on the GPU box, where the reference does not exist.

    python oracle/make_golden.py

Fixtures:
  helm32_c128.npz     helmFE_var(N=32, omega=12, C=1, rho=0.15), b=rhsA(32,12), x0=0,
                      x after 10 / 50 / 200 iterations of helmFE_var.CG      (complex128)
  poisson32_f64.npz   Poisson(32), b=ones, x0=0, x after 10 / 40 / 120 its   (float64)
  helm16_varC.npz     helmFE_var(N=16, omega=7.3, C=U(0.5,1.5), rho=0.21) matrix only
  helm32_pcg.npz      the same helm32 system through the reference's PCG (helmFE_var.py:546-586): no preconditioner
                      and Jacobi (inverse diagonal as a sparse matrix, the `M.dot(r)` branch), tol 1e-4 / 1e-8:
                      x and the iteration index it returned
  clref_poisson32_f32.npz, clref_helm32_c64.npz
                      the reference's OWN OpenCL kernels (oracle/clref: kernel/*/*.cl executed on the CPU inside
                      clcg.c's launch sequence) on the two systems above in single precision: B (k right-hand
                      sides) and x after 10 / 40 iterations for k = 1, after 25 iterations for k = 3
  known_answers.json  iterations to sqrt(|delta_k|/|delta_0|) < tol for the two systems of
                      SURVEY.md 8(c), produced by oracle/np_cg.py (stop=True) after that
                      restatement was checked bit-identical to helmFE_var.CG here.
"""
import ast
import json
import os
import sys

import numpy as np
import scipy
import scipy.sparse

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "conjugate-gradient-pyopencl_b200"))


def lift(path, name, env):
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    mod = ast.Module(body=[fn], type_ignores=[])
    exec(compile(mod, path, "exec"), env)
    return env[name]


def canon(A):
    A = scipy.sparse.csr_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    return A


def main():
    import helmFE_var as H            # the reference module
    import np_cg                      # our restatement
    import problems
    os.makedirs(OUT, exist_ok=True)
    Poisson = lift(os.path.join(REF, "p_helmholtz.py"), "Poisson", {"zeros": np.zeros, "scipy": scipy})

    # --- helm32 (complex128) -------------------------------------------------
    N = 32
    A = canon(H.helmFE_var(N, 12.0, np.ones((N - 1, N - 1)), 0.15, N, N))
    b = H.rhsA(N, 12.0).flatten()
    xs = {}
    for it in (10, 50, 200):
        xs[it] = H.CG(A, b, x=np.zeros(N * N, dtype=complex), maxit=it)
        mine = np_cg.cg(A, b, x=np.zeros(N * N, dtype=complex), maxit=it)
        assert np.array_equal(xs[it], mine), "np_cg is not bit-identical to helmFE_var.CG"
    np.savez_compressed(os.path.join(OUT, "helm32_c128.npz"), data=A.data, indices=A.indices.astype(np.int32),
                        indptr=A.indptr.astype(np.int32), b=b, x10=xs[10], x50=xs[50], x200=xs[200])

    # --- poisson32 (float64) -------------------------------------------------
    N = 32
    A = canon(Poisson(N))
    b = np.ones(N * N)
    xs = {}
    for it in (10, 40, 120):
        xs[it] = H.CG(A, b, x=np.zeros(N * N), maxit=it)
        assert np.array_equal(xs[it], np_cg.cg(A, b, x=np.zeros(N * N), maxit=it))
    np.savez_compressed(os.path.join(OUT, "poisson32_f64.npz"), data=A.data, indices=A.indices.astype(np.int32),
                        indptr=A.indptr.astype(np.int32), b=b, x10=xs[10], x40=xs[40], x120=xs[120])

    # --- variable wave speed matrix ------------------------------------------
    N = 16
    C = np.random.default_rng(16).random((N - 1, N - 1)) + 0.5
    A = canon(H.helmFE_var(N, 7.3, C, 0.21, N, N))
    np.savez_compressed(os.path.join(OUT, "helm16_varC.npz"), data=A.data, indices=A.indices.astype(np.int32),
                        indptr=A.indptr.astype(np.int32), C=C)

    # --- known answers ----------------------------------------
    ka = {}
    A = canon(Poisson(256))
    assert (problems.poisson2d(256) != A).nnz == 0
    b = np.ones(256 * 256)
    ka["poisson256_f64"] = {"n": int(A.shape[0]), "nnz": int(A.nnz), "iters": {}}
    for tol in (1e-6, 1e-8, 1e-10, 1e-12):
        _, it = np_cg.cg(A, b, x=np.zeros(b.size), tol=tol, maxit=2000, stop=True)
        ka["poisson256_f64"]["iters"][f"{tol:g}"] = int(it)
    A = canon(H.helmFE_var(128, 12.0, np.ones((127, 127)), 0.15, 128, 128))
    b = H.rhsA(128, 12.0).flatten()
    ka["helm128_c128"] = {"n": int(A.shape[0]), "nnz": int(A.nnz), "iters": {}}
    for tol in (1e-6, 1e-8, 1e-10):
        x, it = np_cg.cg(A, b, x=np.zeros(b.size, dtype=complex), tol=tol, maxit=3000, stop=True)
        ka["helm128_c128"]["iters"][f"{tol:g}"] = int(it)
    ka["helm128_c128"]["true_relres_at_1e-10"] = float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))
    A = canon(H.helmFE_var(32, 12.0, np.ones((31, 31)), 0.15, 32, 32))
    b = H.rhsA(32, 12.0).flatten()
    ka["helm32_c128"] = {"n": int(A.shape[0]), "nnz": int(A.nnz), "true_relres": {}}
    for it in (10, 50, 200):
        x = H.CG(A, b, x=np.zeros(b.size, dtype=complex), maxit=it)
        ka["helm32_c128"]["true_relres"][str(it)] = float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))
    json.dump(ka, open(os.path.join(OUT, "known_answers.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(ka, indent=1, sort_keys=True))


def pcg_fixture():
    """helm32_pcg.npz -- the reference's PCG, see the header."""
    import warnings
    import helmFE_var as H
    import problems
    A = problems.helmholtz_fe(32)
    b = problems.rhs_a(32, 12.0)
    dinv = 1.0 / A.diagonal()
    Minv = scipy.sparse.csr_matrix(scipy.sparse.diags(dinv))
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # PCG tests `type(M) is scipy.sparse.csr.csr_matrix`: a deprecated path
        for name, M in (("none", None), ("jacobi", Minv)):
            for tol in (1e-4, 1e-8):
                x, i = H.PCG(A, b, M=M, tol=tol, maxit=500)
                out[f"x_{name}_{tol:g}"] = x
                out[f"i_{name}_{tol:g}"] = i
    np.savez_compressed(os.path.join(OUT, "helm32_pcg.npz"), dinv=dinv, **out)


def clref_fixtures():
    """clref_*.npz -- results of the reference's own kernels, see the header."""
    import clref
    import problems
    for name, A, b0, dt in (("poisson32_f32", problems.poisson2d(32), np.ones(1024), np.float32),
                            ("helm32_c64", problems.helmholtz_fe(32), problems.rhs_a(32, 12.0), np.complex64)):
        A = canon(A)
        n = A.shape[0]
        vals = A.data.astype(dt)
        rng = np.random.default_rng(2024)
        extra = [rng.standard_normal(n) + (1j * rng.standard_normal(n) if dt == np.complex64 else 0) for _ in range(2)]
        B3 = np.concatenate([b0] + extra).astype(dt)
        out = {"B3": B3}
        for its in (10, 40):
            out[f"x_k1_it{its}"] = clref.cg(vals, A.indptr, A.indices, B3[:n], k=1, iters=its)
        out["x_k3_it25"] = clref.cg(vals, A.indptr, A.indices, B3, k=3, iters=25)
        np.savez_compressed(os.path.join(OUT, f"clref_{name}.npz"), **out)


def local_rect_fixture():
    """local_rect_9x7.npz / local_rect_12x12.npz -- what the reference's own `local_rect` (p_helmholtz.py:1342-1542, lifted
    with `ast` like Poisson) returns for a non-square and a square subdomain, with the call's parameters: the golden
    arrays of the device-side assembly (conjugate-gradient-pyopencl_b200/assemble.py, csrc/assemble.cuh)."""
    local_rect = lift(os.path.join(REF, "p_helmholtz.py"), "local_rect", {"zeros": np.zeros, "scipy": scipy})
    for tag, (N, k, eps, eta, L, Nh, Nv) in (("9x7", (24, 20.0, 20.0, 20.0, 1.0, 9, 7)),
                                             ("12x12", (40, 12.5, 3.0, 12.5, 2.0, 12, 12))):
        A = canon(local_rect(N, k, eps, eta, L, Nh, Nv))
        np.savez_compressed(os.path.join(OUT, f"local_rect_{tag}.npz"), data=A.data, indices=A.indices.astype(np.int32),
                            indptr=A.indptr.astype(np.int32), params=np.array([N, k, eps, eta, L, Nh, Nv], dtype=np.float64))


if __name__ == "__main__":
    if "--local-rect-only" in sys.argv:
        local_rect_fixture()
        sys.exit(0)
    main()
    pcg_fixture()
    clref_fixtures()
    local_rect_fixture()
