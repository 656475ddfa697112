"""The synthetic systems of the BASELINE.json configs against the reference's generators."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import cg_b200.problems as P


def _golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    n = g["indptr"].size - 1
    return g, sp.csr_matrix((g["data"], g["indices"], g["indptr"]), shape=(n, n))


def test_helmholtz_equals_reference_fixture(golden_dir):
    g, A = _golden(golden_dir, "helm32_c128.npz")
    B = P.helmholtz_fe(32)
    assert np.array_equal(B.indptr, A.indptr) and np.array_equal(B.indices, A.indices)
    assert np.array_equal(B.data, A.data)          # bit for bit
    assert np.array_equal(P.rhs_a(32, 12.0), g["b"])


def test_helmholtz_variable_speed_equals_reference_fixture(golden_dir):
    g, A = _golden(golden_dir, "helm16_varC.npz")
    B = P.helmholtz_fe(16, 7.3, 0.21, g["C"])
    assert np.array_equal(B.indices, A.indices)
    assert np.allclose(B.data, A.data, rtol=1e-14, atol=0)


def test_poisson_equals_reference_fixture(golden_dir):
    g, A = _golden(golden_dir, "poisson32_f64.npz")
    B = P.poisson2d(32)
    assert np.array_equal(B.indptr, A.indptr) and np.array_equal(B.indices, A.indices)
    assert np.array_equal(B.data, A.data)


@pytest.mark.reference
@pytest.mark.parametrize("N", [3, 8, 33, 64])
def test_helmholtz_equals_imported_reference(N):
    import sys
    sys.path.insert(0, "/root/reference")
    import helmFE_var as H
    A = sp.csr_matrix(H.helmFE_var(N, 12.0, np.ones((N - 1, N - 1)), 0.15, N, N))
    A.sum_duplicates()
    A.sort_indices()
    B = P.helmholtz_fe(N)
    assert np.array_equal(B.indices, A.indices) and np.array_equal(B.data, A.data)
    assert np.array_equal(P.rhs_a(N, 12.0), H.rhsA(N, 12.0).flatten())


def test_matrix_properties():
    A = P.helmholtz_fe(64)
    assert A.nnz == 64 ** 2 + 4 * 64 * 63 + 2 * 63 ** 2            # helmFE_var.py:58
    assert abs(A - A.T).max() == 0                                  # complex symmetric
    assert abs(A - A.conj().T).max() > 0.1                          # not Hermitian
    rl = np.diff(A.indptr)
    assert rl.min() == 3 and rl.max() == 7
    A = P.poisson2d(256)
    assert (A.shape[0], A.nnz) == (65536, 326656)
    assert abs(A - A.T).max() == 0 and np.all(A.diagonal() == 4.0)
    A = P.laplace3d(12)
    assert A.nnz == 7 * 12 ** 3 - 6 * 12 ** 2
    assert abs(A - A.T).max() == 0 and np.all(A.diagonal() == 6.0) and A.has_sorted_indices
    assert np.all(np.diff(A.indptr) >= 4)
    # against a Kronecker-sum construction
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(12, 12))
    I = sp.eye(12)
    K = sp.kron(sp.kron(T, I), I) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(I, I), T)
    assert abs(A - K).max() == 0


def test_powerlaw_is_spd_and_heavy_tailed():
    A = P.powerlaw_spd(n=20000, nnz_target=200000)
    assert abs(A - A.T).max() == 0
    d = A.diagonal()
    off = np.asarray(abs(A).sum(axis=1)).ravel() - d
    assert np.all(d > off)                                          # strictly diagonally dominant
    rl = np.diff(A.indptr)
    assert 0.8 * 200000 < A.nnz < 1.1 * 200000
    assert rl.max() > 20 * rl.mean()
    assert A.indptr.dtype == np.int32 and A.indices.dtype == np.int32


def test_byte_model():
    # B_spmv 184.4 MB, B_iter 335.4 MB
    bs, bi = P.algorithmic_bytes(1048576, 7331842, 1, "c128")
    assert abs(bs / 1e6 - 184.4) < 0.1 and abs(bi / 1e6 - 335.4) < 0.1
    bs, bi = P.algorithmic_bytes(2097152, 14581760, 32, "f64")
    assert abs(bs / 1e9 - 1.257) < 0.001 and abs(bi / 1e9 - 6.089) < 0.001


def test_constant_coefficient_grid_operators_have_a_handful_of_distinct_rows(golden_dir):
    """Why the row-pattern dictionary applies to the reference's own workload: as lists of
    (column - row, value) the rows of the BASELINE grid operators, and of the subdomain matrices the unmodified
    driver built (`local_rect`, recorded in tests/golden/asprec_*.npz), come in 9 (2-D) or 27 (3-D) kinds --
    in double and after the drivers' cast to csingle alike; a power-law matrix has as many kinds as rows."""
    import os
    import scipy.sparse as sp
    import cg_b200.problems as P
    assert P.row_patterns(P.poisson2d(40)) == 9
    assert P.row_patterns(P.laplace3d(12)) == 27
    assert P.row_patterns(P.helmholtz_fe(48)) == 9
    for name in ("asprec_2_12.npz", "asprec_3_16.npz"):
        z = np.load(os.path.join(golden_dir, name))
        n = int(z["n"])
        A = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=(n, n))
        assert P.row_patterns(A) == 9 and P.row_patterns(A.astype(np.complex64)) == 9
    R = P.powerlaw_spd(n=3000, nnz_target=30000, max_row=400)
    assert P.row_patterns(R) == R.shape[0]


# ---------------------------------------------------------------------------------------
# class tables of the device-side assembly (conjugate-gradient-pyopencl_b200/assemble.py): the numpy expansion of a table
# is the checker of csrc/assemble.cuh, so it is pinned here against the reference's own generators
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["9x7", "12x12"])
def test_local_rect_class_table_reproduces_the_reference_fixture_bit_for_bit(golden_dir, tag):
    """tests/golden/local_rect_*.npz: what the reference's `local_rect` (p_helmholtz.py:1342-1542) returned in the build
    container (oracle/make_golden.py::local_rect_fixture)."""
    import cg_b200.problems as P
    z = np.load(os.path.join(golden_dir, f"local_rect_{tag}.npz"))
    N, k, eps, eta, L, Nh, Nv = z["params"]
    A = P.local_rect(N, k, eps, eta, L, int(Nh), int(Nv))
    assert np.array_equal(A.indptr, z["indptr"]) and np.array_equal(A.indices, z["indices"])
    assert np.array_equal(A.data.view(np.float64), z["data"].view(np.float64))
    assert abs(A - A.T).max() == 0 and P.row_patterns(A) == 9              # complex symmetric, one pattern per node class


@pytest.mark.reference
def test_local_rect_class_table_against_the_reference_function():
    import ast
    import scipy
    import cg_b200.problems as P
    path = os.path.join("/root/reference", "p_helmholtz.py")
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "local_rect")
    env = {"zeros": np.zeros, "scipy": scipy}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), env)
    rng = np.random.default_rng(5)
    for Nh, Nv in ((4, 4), (5, 9), (16, 3), (11, 11)):
        N, k, eps, eta, L = int(rng.integers(8, 60)), float(rng.uniform(1, 40)), float(rng.uniform(0, 30)), float(rng.uniform(1, 40)), float(rng.uniform(0.5, 2))
        R = scipy.sparse.csr_matrix(env["local_rect"](N, k, eps, eta, L, Nh, Nv))
        R.sum_duplicates()
        R.sort_indices()
        A = P.local_rect(N, k, eps, eta, L, Nh, Nv)
        assert np.array_equal(A.indptr, R.indptr) and np.array_equal(A.indices, R.indices)
        assert np.array_equal(A.data.view(np.float64), R.data.view(np.float64))


def test_class_tables_of_the_other_operators_expand_to_the_host_generators():
    import cg_b200.problems as P
    from cg_b200 import assemble
    same = lambda A, B: (np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
                         and np.array_equal(np.ascontiguousarray(A.data).view(np.float64), np.ascontiguousarray(B.data).view(np.float64)))
    for N in (4, 7, 33):
        assert same(assemble.expand(assemble.poisson2d_table(), N, N, 1, np.float64), P.poisson2d(N))
        T = assemble.table_from_template(P.helmholtz_fe(5, h=1.0 / (N - 1.0)), (5, 5))
        assert same(assemble.expand(T, N, N, 1, np.complex128), P.helmholtz_fe(N))
    assert same(assemble.expand(assemble.laplace3d_table(), 6, 6, 5, np.float64), P.laplace3d(6, nz=5))
    assert same(assemble.expand(assemble.laplace3d_table(), 4, 4, 4, np.float64), P.laplace3d(4))
    # grids with a direction of 2 or 3 points (no interior / one interior node)
    assert same(assemble.expand(assemble.poisson2d_table(), 2, 2, 1, np.float64),
                assemble.expand(assemble.table_from_template(P.poisson2d(4), (4, 4)), 2, 2, 1, np.float64))
