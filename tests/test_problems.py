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
