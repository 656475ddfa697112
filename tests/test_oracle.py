"""Pins the two CPU oracles (test infrastructure) before anything is compared with them.

  - oracle/np_cg.py  against fixtures produced by the reference's own numpy CG
    (helmFE_var.py:507-544) -- bit-identical;
  - oracle/cpu_ref.c (restatement of clcg.c:253-419 + kernels) against the same fixtures
    and against the known answers of SURVEY.md 8(c) (tests/golden/known_answers.json).
"""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

import np_cg


def _load(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    n = g["indptr"].size - 1
    A = sp.csr_matrix((g["data"], g["indices"], g["indptr"]), shape=(n, n))
    return g, A


@pytest.mark.parametrize("name,its,cplx", [("helm32_c128.npz", (10, 50, 200), True),
                                           ("poisson32_f64.npz", (10, 40, 120), False)])
def test_np_cg_bit_identical_to_reference_fixture(golden_dir, name, its, cplx):
    g, A = _load(golden_dir, name)
    for it in its:
        x0 = np.zeros(A.shape[0], dtype=complex if cplx else float)
        x = np_cg.cg(A, g["b"], x=x0, maxit=it)
        assert np.array_equal(x, g[f"x{it}"])


@pytest.mark.reference
def test_np_cg_bit_identical_to_imported_reference():
    import sys
    sys.path.insert(0, "/root/reference")
    import helmFE_var as H
    N = 24
    A = H.helmFE_var(N, 12.0, np.ones((N - 1, N - 1)), 0.15, N, N)
    b = H.rhs(N, 12.0).flatten()
    for it in (1, 7, 60):
        assert np.array_equal(H.CG(A, b, x=np.zeros(N * N, dtype=complex), maxit=it),
                              np_cg.cg(A, b, x=np.zeros(N * N, dtype=complex), maxit=it))


@pytest.mark.parametrize("name,its", [("helm32_c128.npz", (10, 50, 200)), ("poisson32_f64.npz", (10, 40, 120))])
def test_c_oracle_matches_reference_fixture_double(cpu_ref, golden_dir, name, its):
    g, A = _load(golden_dir, name)
    for it in its:
        x, n_it, _ = cpu_ref.cg(g["data"], g["indptr"], g["indices"], g["b"], iters=it)
        ref = g[f"x{it}"]
        # different summation order (GPU-style trees) => agreement to rounding, not bitwise
        assert np.linalg.norm(x - ref) <= 1e-12 * np.linalg.norm(ref)
        assert n_it[0] == it


def test_c_oracle_single_precision_tracks_double(cpu_ref, golden_dir):
    g, A = _load(golden_dir, "helm32_c128.npz")
    x, _, _ = cpu_ref.cg(g["data"].astype(np.complex64), g["indptr"], g["indices"],
                         g["b"].astype(np.complex64), iters=50)
    ref = g["x50"]
    assert np.linalg.norm(x - ref) <= 2e-4 * np.linalg.norm(ref)


def test_c_oracle_known_answers(cpu_ref, golden_dir):
    import cg_b200.problems as P
    ka = json.load(open(os.path.join(golden_dir, "known_answers.json")))
    A = P.poisson2d(256)
    assert (A.shape[0], A.nnz) == (ka["poisson256_f64"]["n"], ka["poisson256_f64"]["nnz"])
    b = np.ones(A.shape[0])
    for tol, want in ka["poisson256_f64"]["iters"].items():
        _, its, _ = cpu_ref.cg(A.data, A.indptr, A.indices, b, iters=2000, tol=float(tol))
        assert its[0] == want
    A = P.helmholtz_fe(128)
    assert (A.shape[0], A.nnz) == (ka["helm128_c128"]["n"], ka["helm128_c128"]["nnz"])
    b = P.rhs_a(128, 12.0)
    for tol, want in ka["helm128_c128"]["iters"].items():
        x, its, _ = cpu_ref.cg(A.data, A.indptr, A.indices, b, iters=3000, tol=float(tol))
        assert its[0] == want
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) < 1e-9
    # single precision, SURVEY.md 8(c): 271 / 328 iterations at 1e-4 / 1e-5
    a32, b32 = A.data.astype(np.complex64), b.astype(np.complex64)
    for tol, want in ((1e-4, 271), (1e-5, 328)):
        _, its, _ = cpu_ref.cg(a32, A.indptr, A.indices, b32, iters=3000, tol=tol)
        assert abs(int(its[0]) - want) <= 1


def test_c_oracle_delta_history_and_np_cg_agree(cpu_ref, golden_dir):
    g, A = _load(golden_dir, "helm32_c128.npz")
    hist = []
    np_cg.cg(A, g["b"], x=np.zeros(A.shape[0], dtype=complex), maxit=60, history=hist)
    _, _, h = cpu_ref.cg(g["data"], g["indptr"], g["indices"], g["b"], iters=60, want_hist=True)
    hist = np.array(hist)
    assert h.shape == (61, 1)
    assert np.all(np.abs(h[:, 0] - hist) <= 1e-10 * np.abs(hist))


def test_c_oracle_multi_rhs_columns_are_independent(cpu_ref, golden_dir):
    g, A = _load(golden_dir, "poisson32_f64.npz")
    n = A.shape[0]
    rng = np.random.default_rng(3)
    B = rng.standard_normal((3, n))
    X, _, _ = cpu_ref.cg(g["data"], g["indptr"], g["indices"], B.ravel(), k=3, iters=30)
    for r in range(3):
        x1, _, _ = cpu_ref.cg(g["data"], g["indptr"], g["indices"], B[r], iters=30)
        assert np.array_equal(X[r * n:(r + 1) * n], x1)


def test_c_oracle_edge_shapes(cpu_ref):
    # n < 256 and n % 8 != 0 (the reference's "NOT SUPPORTED" / out-of-bounds cases), rows longer than a wave
    rng = np.random.default_rng(0)
    for n in (1, 5, 37, 255, 257, 300):
        M = sp.random(n, n, density=min(1.0, 40.0 / n), random_state=1, format="csr")
        A = (M + M.T + sp.eye(n) * (n + 1.0)).tocsr()
        A.sort_indices()
        b = rng.standard_normal(n)
        x, _, _ = cpu_ref.cg(A.data, A.indptr, A.indices, b, iters=min(n, 40) + 5, tol=1e-13)
        assert np.linalg.norm(A @ x - b) <= 1e-10 * np.linalg.norm(b)
        y = cpu_ref.spmv(A.data, A.indptr, A.indices, b)
        assert np.allclose(y, A @ b, rtol=1e-13, atol=1e-13)


# ---------------------------------------------------------------------------------------
# the reference driver's own subdomain solves (oracle/run_reference_driver.py ran p_h-PY_C-CL.py unmodified)
# ---------------------------------------------------------------------------------------
ASPREC = ["asprec_2_12.npz", "asprec_3_16.npz"]


def _asprec(golden_dir, name):
    import scipy.sparse as sp
    z = np.load(os.path.join(golden_dir, name))
    n = int(z["n"])
    return z, sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=(n, n))


@pytest.mark.parametrize("name", ASPREC)
def test_driver_numpy_cg_restatement_is_bit_identical_to_the_reference_run(golden_dir, name):
    """x_numpy_cg is what the reference driver's CG (p_h-PY_C-CL.py:1338-1369) returned inside as_prec when the
    driver itself ran here; oracle/np_cg.cg_abs_tol must reproduce every bit, and the iteration counts."""
    import np_cg
    z, A = _asprec(golden_dir, name)
    assert (abs(A - A.T) > 0).nnz == 0 and abs(A - A.conj().T).max() > 0.1      # complex-symmetric, not Hermitian
    for p in range(z["z"].shape[0]):
        x, it = np_cg.cg_abs_tol(A, z["z"][p], tol=float(z["numpy_cg_tol"]))
        assert np.array_equal(x, z["x_numpy_cg"][p])
        assert it == int(z["numpy_cg_iters"][p])


@pytest.mark.parametrize("name", ASPREC)
def test_c_oracle_on_the_driver_subdomain_systems(cpu_ref, golden_dir, name):
    """The C restatement (device summation order) against the reference-run result at the same iteration count.
    These subdomain operators are indefinite; after ~100 COCG iterations two summation orders of the same
    double arithmetic are 4e-13 .. 6e-9 apart (measured), so the bar here is 1e-7, not 1e-10."""
    z, A = _asprec(golden_dir, name)
    for p in range(z["z"].shape[0]):
        x, _, _ = cpu_ref.cg(A.data, A.indptr, A.indices, z["z"][p], iters=int(z["numpy_cg_iters"][p]))
        ref = z["x_numpy_cg"][p]
        assert np.linalg.norm(x - ref) / np.linalg.norm(ref) < 1e-7


@pytest.mark.parametrize("name", ASPREC)
def test_as_prec_builds_the_block_the_abi_expects(golden_dir, name):
    """What as_prec hands to pcl.CG in its multi-RHS variant (p_h-PY_C-CL.py:1924-1937): csingle / intc arrays,
    RHS p at offset p*size, x0 = 0, the SAME matrix for every subdomain."""
    z, A = _asprec(golden_dir, name)
    k, size = int(z["cl_args_n_rhs"]), int(z["cl_args_size"])
    assert size == A.shape[0] and int(z["cl_args_nnz"]) == A.nnz and k == z["z"].shape[0]
    assert z["cl_args_a_values"].dtype == np.csingle and z["cl_args_b_values"].dtype == np.csingle
    assert z["cl_args_a_pointers"].dtype == np.intc and z["cl_args_a_cols"].dtype == np.intc
    assert np.array_equal(z["cl_args_a_values"], A.data.astype(np.csingle))
    assert np.array_equal(z["cl_args_b_values"].reshape(k, size), z["z"].astype(np.csingle))
    assert not z["cl_args_x_in"].any()
    # ... and what the reference's own kernels return for exactly these arrays is what the C oracle returns
    import cpu_ref as _cpu_ref
    _cpu_ref.build()
    mine, _, _ = _cpu_ref.cg(z["cl_args_a_values"], z["cl_args_a_pointers"], z["cl_args_a_cols"], z["cl_args_b_values"],
                             x0=z["cl_args_x_in"], k=k, iters=int(z["cl_args_n_iterations"]))
    assert np.array_equal(mine.view(np.uint8), z["cl_result_reference_kernels"].view(np.uint8))
    g = z["gmres_iterations"]                   # variants 0 (exact), 1, 2 (CGMaxIT fixed, single), 5 (numpy CG to 1e-5)
    assert len(g) == 4 and g[0] == g[3] and g[1] == g[2] >= g[0]


@pytest.mark.reference
def test_reference_driver_still_produces_the_committed_fixture(golden_dir, tmp_path):
    """Build container only: run the unmodified driver again and compare with the committed fixture."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    old = dict(np.load(os.path.join(golden_dir, "asprec_2_12.npz")))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    subprocess.check_call([sys.executable, os.path.join(root, "oracle", "run_reference_driver.py"), "2", "12", "40",
                           "--out", str(tmp_path)], env=env, stdout=subprocess.DEVNULL)
    new = np.load(os.path.join(str(tmp_path), "asprec_2_12.npz"))
    for key in ("data", "indices", "indptr", "z", "x_numpy_cg", "numpy_cg_iters", "cl_args_b_values", "gmres_iterations",
                "cl_result_reference_kernels"):
        assert np.array_equal(old[key], new[key]), key


@pytest.mark.parametrize("name", ["none", "jacobi"])
@pytest.mark.parametrize("tol", [1e-4, 1e-8])
def test_pcg_restatement_is_bit_identical_to_the_reference_pcg(golden_dir, name, tol):
    """oracle/np_cg.pcg against what the reference's own PCG (helmFE_var.py:546-586) returned here for the helm32
    system -- the oracle for the preconditioned solver of SURVEY.md 8(f) rank 2 (not built yet)."""
    import cg_b200.problems as P
    z = np.load(os.path.join(golden_dir, "helm32_pcg.npz"))
    A, b = P.helmholtz_fe(32), P.rhs_a(32, 12.0)
    x, i = np_cg.pcg(A, b, M=None if name == "none" else z["dinv"], tol=tol, maxit=500)
    assert i == int(z[f"i_{name}_{tol:g}"])
    assert np.array_equal(x, z[f"x_{name}_{tol:g}"])
    assert int(z["i_jacobi_1e-08"]) < int(z["i_none_1e-08"])           # Jacobi helps a little on this operator


@pytest.mark.reference
def test_reference_multi_gpu_script_splits_right_hand_sides_over_our_cl_module():
    """Build container only: p_h-PY_C-CL-multi-GPU.py, unmodified, on two pretend devices.  Its own
    distribute_workloads_on_devices / distribute_computations_with_threads (:2123-2181) drive this repo's
    cl.conjugate_gradient_multi_gpu from one thread per device; the stitched result equals the single multi-RHS
    call, and the ranges are what sharded.split_rhs computes."""
    import ast
    import subprocess
    import sys
    from cg_b200 import sharded
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.check_output([sys.executable, os.path.join(root, "oracle", "run_reference_driver.py"), "--multi-gpu"],
                                  text=True, env=dict(os.environ, OMP_NUM_THREADS="2"))
    res = ast.literal_eval(out.strip().splitlines()[-1])
    assert res["identical"] is True
    assert res["devices"] == [0, 1] and res["split_calls"] == [4, 5] and res["n_my"] == 9
    assert [tuple(r) for r in res["ranges"]] == [tuple(r) for r in sharded.split_rhs(9, 2)]


@pytest.mark.reference
@pytest.mark.parametrize("use_cg,n_rhs,calls", [(1, [1], 28), (2, [4], 7)])
def test_oldest_reference_driver_runs_on_the_old_cl_api(use_cg, n_rhs, calls):
    """Build container only: p_helmholtz.py, unmodified -- `pcl.create_kernels(1)` at import and the 9-argument
    `pcl.CG(size, nnz, a, b, ptr, cols, x, n_rhs, maxit)` (:31, :1839, :1873) -- runs to the end of its GMRES
    (7 iterations, as with exact subdomain solves) on this repo's cl.py."""
    import ast
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.check_output([sys.executable, os.path.join(root, "oracle", "run_reference_driver.py"), "--old-api",
                                   str(use_cg)], text=True, env=dict(os.environ, OMP_NUM_THREADS="2"))
    res = ast.literal_eval(out.strip().splitlines()[-1])
    assert res["n_rhs"] == n_rhs and res["cl_calls"] == calls and res["gmres_iterations"] == [7]


# ---------------------------------------------------------------------------------------
# the reference's own OpenCL kernels (oracle/clref: kernel/*/*.cl from /root/reference, executed on the CPU)
# ---------------------------------------------------------------------------------------
def _clref_system(kind):
    import cg_b200.problems as P
    if kind == "poisson32_f32":
        return P.poisson2d(32), np.float32
    return P.helmholtz_fe(32), np.complex64


@pytest.mark.reference
@pytest.mark.parametrize("kind", ["poisson32_f32", "helm32_c64"])
@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_c_oracle_is_bit_identical_to_the_reference_kernels(cpu_ref, kind, k):
    """Build container only.  oracle/cpu_ref.c RESTATES the reference's kernels; oracle/clref EXECUTES them (the .cl
    sources #included from /root/reference, work-items as fibers, barriers honoured, clcg.c's launch sequence and
    host sums around them).  Same bits after 0, 1 and 12 iterations for 1..4 right-hand sides, real and complex --
    the summation trees, the host partial sums and the complex arithmetic of the restatement are the reference's."""
    import clref
    A, dt = _clref_system(kind)
    A.sort_indices()
    n = A.shape[0]
    rng = np.random.default_rng(100 + k)
    B = np.concatenate([rng.standard_normal(n) + (1j * rng.standard_normal(n) if dt == np.complex64 else 0)
                        for _ in range(k)]).astype(dt)
    X0 = (0.1 * np.concatenate([rng.standard_normal(n) + (1j * rng.standard_normal(n) if dt == np.complex64 else 0)
                                for _ in range(k)])).astype(dt)
    vals = A.data.astype(dt)
    for its, x0 in ((0, X0), (1, None), (12, X0)):
        ref = clref.cg(vals, A.indptr, A.indices, B, x0=x0, k=k, iters=its)
        mine, _, _ = cpu_ref.cg(vals, A.indptr, A.indices, B, x0=x0, k=k, iters=its)
        assert np.array_equal(ref.view(np.uint8), mine.view(np.uint8)), (kind, k, its)


@pytest.mark.parametrize("kind", ["poisson32_f32", "helm32_c64"])
def test_c_oracle_reproduces_the_reference_kernel_fixtures(cpu_ref, golden_dir, kind):
    """Runs everywhere (the GPU box included): tests/golden/clref_*.npz are results of the reference's own kernels
    (oracle/make_golden.py::clref_fixtures); the C oracle must give the same bits."""
    z = np.load(os.path.join(golden_dir, f"clref_{kind}.npz"))
    A, dt = _clref_system(kind)
    A.sort_indices()
    n = A.shape[0]
    vals = A.data.astype(dt)
    for its in (10, 40):
        x, _, _ = cpu_ref.cg(vals, A.indptr, A.indices, z["B3"][:n], k=1, iters=its)
        assert np.array_equal(x.view(np.uint8), z[f"x_k1_it{its}"].view(np.uint8)), its
    x, _, _ = cpu_ref.cg(vals, A.indptr, A.indices, z["B3"], k=3, iters=25)
    assert np.array_equal(x.view(np.uint8), z["x_k3_it25"].view(np.uint8))


# ---------------------------------------------------------------------------------------
# randomised cross-checks of the oracles (hypothesis)
# ---------------------------------------------------------------------------------------
def _random_spd(n, density, seed, cplx):
    """Diagonally dominant symmetric (complex-symmetric when cplx) CSR matrix with unsorted duplicates removed."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    L = sp.random(n, n, density=density, random_state=int(seed), format="csr")
    L.data = rng.uniform(-1.0, 1.0, L.nnz) + (1j * rng.uniform(-1.0, 1.0, L.nnz) if cplx else 0)
    S = L + L.T
    d = np.asarray(abs(S).sum(axis=1)).ravel() + 1.0
    A = (S + sp.diags(d + (0.2j if cplx else 0))).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def test_randomised_double_oracles_agree():
    """cpu_ref.c in double (device summation order) against np_cg (numpy order) on random diagonally dominant
    systems, real and complex-symmetric, several right-hand sides: two orders of the same recurrence, 1e-10."""
    from hypothesis import given, settings, strategies as st
    import cpu_ref as cr
    cr.build()

    @settings(max_examples=12, deadline=None)
    @given(n=st.integers(256, 700), k=st.integers(1, 4), seed=st.integers(0, 2**31 - 1), cplx=st.booleans(),
           its=st.integers(1, 25))
    def run(n, k, seed, cplx, its):
        A = _random_spd(n, 6.0 / n, seed, cplx)
        rng = np.random.default_rng(seed + 1)
        dt = np.complex128 if cplx else np.float64
        B = np.concatenate([rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0) for _ in range(k)]).astype(dt)
        x, _, _ = cr.cg(A.data.astype(dt), A.indptr, A.indices, B, k=k, iters=its)
        for c in range(k):
            ref = np_cg.cg(A.astype(dt), B[c * n:(c + 1) * n].astype(complex), maxit=its)
            assert np.linalg.norm(x[c * n:(c + 1) * n] - ref) <= 1e-10 * np.linalg.norm(ref)

    run()


@pytest.mark.reference
def test_randomised_single_oracle_is_bit_identical_to_the_reference_kernels():
    """Build container only: random systems (row lengths 1..40, n not a multiple of 8 or 256, 1..4 right-hand
    sides, real and complex) through the reference's own kernels (oracle/clref) and through oracle/cpu_ref.c."""
    from hypothesis import given, settings, strategies as st
    import clref
    import cpu_ref as cr
    cr.build()

    @settings(max_examples=10, deadline=None)
    @given(n=st.integers(256, 520), k=st.integers(1, 4), seed=st.integers(0, 2**31 - 1), cplx=st.booleans(),
           its=st.integers(0, 6), dens=st.floats(2.0, 40.0))
    def run(n, k, seed, cplx, its, dens):
        A = _random_spd(n, dens / n, seed, cplx)
        rng = np.random.default_rng(seed + 1)
        dt = np.complex64 if cplx else np.float32
        B = np.concatenate([rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0) for _ in range(k)]).astype(dt)
        vals = A.data.astype(dt)
        ref = clref.cg(vals, A.indptr, A.indices, B, k=k, iters=its)
        mine, _, _ = cr.cg(vals, A.indptr, A.indices, B, k=k, iters=its)
        assert np.array_equal(ref.view(np.uint8), mine.view(np.uint8))

    run()


def test_c_pcg_oracle_against_the_reference_pcg_fixture(cpu_ref, golden_dir):
    """oracle/cpu_ref.c::cpu_ref_pcg (Jacobi, device summation order) after exactly the iterations the reference's
    PCG needed (helm32_pcg.npz: results of helmFE_var.PCG itself): same x to 1e-9 (two summation orders of an
    indefinite problem), and its r.r history crosses the reference's tolerance at the reference's iteration."""
    import cg_b200.problems as P
    z = np.load(os.path.join(golden_dir, "helm32_pcg.npz"))
    A, b = P.helmholtz_fe(32), P.rhs_a(32, 12.0)
    for tol in (1e-4, 1e-8):
        i_ref = int(z[f"i_jacobi_{tol:g}"])
        x, hist = cpu_ref.pcg(A.data, A.indptr, A.indices, b, z["dinv"], iters=i_ref + 1, want_hist=True)
        ref = z[f"x_jacobi_{tol:g}"]
        assert np.linalg.norm(x - ref) <= 1e-9 * np.linalg.norm(ref)
        res = np.sqrt(np.abs(hist[:, 0]))
        assert res[i_ref + 1] < tol <= res[i_ref]
    # without a preconditioner (dinv = 1) it is the plain CG oracle, bit for bit
    x1, _ = cpu_ref.pcg(A.data, A.indptr, A.indices, b, np.ones(A.shape[0]), iters=30)
    x0, _, _ = cpu_ref.cg(A.data, A.indptr, A.indices, b, iters=30)
    assert np.array_equal(x1, x0)
