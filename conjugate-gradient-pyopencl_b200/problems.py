"""Synthetic CSR systems for the five BASELINE.json configs (inputs only).

The reference builds its matrices with O(n) Python loops
(`Poisson`, /root/reference/p_helmholtz.py:1545-1585; `helmFE_var`,
/root/reference/helmFE_var.py:9-331).  These are vectorised generators of the
same discretisations; tests/test_problems.py checks them entry-for-entry against
the reference functions (when /root/reference is mounted) and against the
fixtures in tests/golden/.

All functions return canonical scipy CSR (sorted indices, no duplicates,
int32 index arrays) -- the exact arrays `cg()` takes.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

__all__ = [
    "poisson2d", "helmholtz_fe", "local_rect", "rhs_a", "laplace3d", "powerlaw_spd",
    "csr_arrays", "algorithmic_bytes", "flops_per_iteration", "DTYPES",
]

# dtype name -> (numpy dtype, code in include/cgb200.h, bytes per value, is complex)
DTYPES = {
    "f32": (np.float32, 0, 4, False),
    "f64": (np.float64, 1, 8, False),
    "c64": (np.complex64, 2, 8, True),
    "c128": (np.complex128, 3, 16, True),
}


def _canonical(A: sp.csr_matrix) -> sp.csr_matrix:
    A = sp.csr_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    A.indptr = A.indptr.astype(np.int32)
    A.indices = A.indices.astype(np.int32)
    return A


def poisson2d(N: int) -> sp.csr_matrix:
    """5-point Laplacian on an N x N grid, diag 4 / off -1, node (i, j) -> i*N + j.

    Same matrix as the reference's `Poisson(N)` (p_helmholtz.py:1545-1585)."""
    idx = np.arange(N * N, dtype=np.int64).reshape(N, N)
    rows = [idx.ravel()]
    cols = [idx.ravel()]
    vals = [np.full(N * N, 4.0)]
    for a, b in ((idx[:, 1:], idx[:, :-1]), (idx[:, :-1], idx[:, 1:]),
                 (idx[1:, :], idx[:-1, :]), (idx[:-1, :], idx[1:, :])):
        rows.append(a.ravel())
        cols.append(b.ravel())
        vals.append(np.full(a.size, -1.0))
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(N * N, N * N))
    return _canonical(A.tocsr())


def helmholtz_fe(N: int, omega: float = 12.0, rho: float = 0.15, C=None, h=None) -> sp.csr_matrix:
    """P1 finite-element Helmholtz matrix S = K - (1+i rho) M_k - i B_k on an N x N grid.

    Same discretisation as `helmFE_var(N, omega, C, rho, N, N)`
    (helmFE_var.py:9-331): unit square, two triangles per cell split along the
    SW-NE diagonal, impedance boundary, wave number k = omega / C[cell] per cell
    (`C` has shape (N-1, N-1), default ones).  Complex symmetric, NOT Hermitian.
    Node (j, m) = (x index, y index) is row m*N + j.

    The reference enumerates node classes (corner / edge / interior) with one
    hand-written formula each; here every entry is the same sum over the cells
    that touch it, with absent cells contributing 0, evaluated in real arithmetic
    in the reference's operation order so the values agree to the last bit.
    """
    if C is None:
        C = np.ones((N - 1, N - 1))
    C = np.asarray(C, dtype=np.float64)
    if h is None:
        h = 1.0 / (N - 1.0)                        # (an explicit h: a small instance with the mesh width of a big grid,
    h2 = h ** 2                                    #  from which assemble.helmholtz_fe reads the row classes)
    k = omega / C                                  # (m, j): cell with SW node (j, m)
    kk = np.zeros((N + 1, N + 1))                  # k^2, zero-padded ring: kk[m+1, j+1]
    kk[1:N, 1:N] = k * k
    kp = np.zeros((N + 1, N + 1))                  # k, same padding
    kp[1:N, 1:N] = k
    # the four cells around node (j, m):   nw=(j-1,m) ne=(j,m) sw=(j-1,m-1) se=(j,m-1)
    ne2, nw2, se2, sw2 = kk[1:, 1:], kk[1:, :-1], kk[:-1, 1:], kk[:-1, :-1]
    ne, nw, se, sw = kp[1:, 1:], kp[1:, :-1], kp[:-1, 1:], kp[:-1, :-1]
    has = np.zeros((N + 1, N + 1))
    has[1:N, 1:N] = 1.0
    cne, cnw, cse, csw = has[1:, 1:], has[1:, :-1], has[:-1, 1:], has[:-1, :-1]

    node = np.arange(N * N, dtype=np.int64).reshape(N, N)   # node[m, j]
    mm, jj = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    on_bot, on_top, on_lft, on_rgt = mm == 0, mm == N - 1, jj == 0, jj == N - 1

    def cplx(re, im):
        out = np.empty(re.shape, dtype=np.complex128)
        out.real = re
        out.imag = im
        return out

    rows, cols, vals = [], [], []

    def emit(mask, dm, dj, re, im):
        r = node[mask]
        rows.append(r)
        cols.append(r + dm * N + dj)
        vals.append(cplx(re[mask], im[mask]))

    # diagonal: stiffness = number of adjacent cells; mass (nw + 2 sw + 2 ne + se) h^2/12;
    # boundary mass: every boundary edge at the node gives k_cell * h/3
    S = nw2 + 2.0 * sw2 + 2.0 * ne2 + se2
    stiff = cne + cnw + cse + csw
    bsum = (np.where(on_bot, nw + ne, 0.0) + np.where(on_top, sw + se, 0.0)
            + np.where(on_lft & ~(on_bot | on_top), ne + se, 0.0)
            + np.where(on_rgt & ~(on_bot | on_top), nw + sw, 0.0))
    # corners: both boundary edges belong to the single adjacent cell -> k*2
    corner = (on_bot | on_top) & (on_lft | on_rgt)
    kc = ne + nw + se + sw                                   # the only non-zero one at a corner
    bsum = np.where(corner, kc * 2, bsum)
    d_re = stiff - S * h2 / 12.0
    d_im = -(rho * S * h2 / 12.0) - bsum * h / 3.0
    # the reference writes the top-right corner as k^2*(h2/6.) (helmFE_var.py:104), which
    # rounds differently from k^2*h2/6.; keep its grouping so the matrices agree bit for bit
    k2 = sw2[N - 1, N - 1]
    d_re[N - 1, N - 1] = 1.0 - k2 * (h2 / 6.0)
    d_im[N - 1, N - 1] = -(rho * k2 * (h2 / 6.0)) - bsum[N - 1, N - 1] * h / 3.0
    emit(np.ones((N, N), bool), 0, 0, d_re, d_im)

    # axis neighbours: -0.5 per cell sharing the edge, mass (two cells) h^2/24,
    # boundary edge: - i k h/6
    def axis(mask, dm, dj, c1, c2, k1sq, k2sq, kedge):
        S2 = k1sq + k2sq
        emit(mask, dm, dj, -0.5 * (c1 + c2) - S2 * h2 / 24.0,
             -(rho * S2 * h2 / 24.0) - kedge * h / 6.0)

    horiz_edge = np.where(on_bot | on_top, 1.0, 0.0)
    vert_edge = np.where(on_lft | on_rgt, 1.0, 0.0)
    axis(~on_rgt, 0, +1, cne, cse, ne2, se2, horiz_edge * (ne + se))    # east  (one of ne/se is 0 on an edge)
    axis(~on_lft, 0, -1, cnw, csw, nw2, sw2, horiz_edge * (nw + sw))    # west
    axis(~on_top, +1, 0, cnw, cne, nw2, ne2, vert_edge * (nw + ne))     # north
    axis(~on_bot, -1, 0, csw, cse, sw2, se2, vert_edge * (sw + se))     # south
    # diagonal neighbours along the cell split: NE through cell ne, SW through cell sw
    emit(~on_rgt & ~on_top, +1, +1, -(ne2 * h2 / 12.0), -(rho * ne2 * h2 / 12.0))
    emit(~on_lft & ~on_bot, -1, -1, -(sw2 * h2 / 12.0), -(rho * sw2 * h2 / 12.0))

    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(N * N, N * N))
    return _canonical(A.tocsr())


def local_rect(N, k, eps, eta, L, Nhoriz, Nvert) -> sp.csr_matrix:
    """The subdomain operator `local_rect(N, k, eps, eta, L, Nhoriz, Nvert)` of the reference's drivers
    (p_helmholtz.py:1342-1542) on the host: the class table of assemble.local_rect_table expanded with numpy."""
    try:
        from . import assemble
    except ImportError:
        import assemble
    return _canonical(assemble.expand(assemble.local_rect_table(N, k, eps, eta, L), Nhoriz, Nvert, 1, np.complex128))


def rhs_a(N: int, omega: float) -> np.ndarray:
    """`rhsA(N, k).flatten()` of the reference (helmFE_var.py:379-389): k^2 on the boundary ring."""
    b = np.zeros((N, N), dtype=np.complex128)
    b[:, 0] = b[:, -1] = b[0, :] = b[-1, :] = omega * omega
    return b.ravel()


def laplace3d(N: int, dtype=np.float64, rows=None, nz=None) -> sp.csr_matrix:
    """7-point Laplacian on an N^3 grid (N x N x nz when nz is given), diag 6 / off -1, Dirichlet
    truncation, x fastest.

    The 3-D extension of `Poisson` named by BASELINE.json configs 3 and 4 (not in the
    reference).  Built straight into CSR arrays: 300^3 has 188 M non-zeros.
    rows=(r0, r1): only that block of rows (global column indices, shape (r1-r0, N^3)) -- what one
    rank of a row-block sharded run needs."""
    NZ = N if nz is None else int(nz)
    n = N * N * NZ
    r0, r1 = (0, n) if rows is None else (int(rows[0]), int(rows[1]))
    idx = np.arange(r0, r1, dtype=np.int64)
    x = (idx % N).astype(np.int32)
    y = ((idx // N) % N).astype(np.int32)
    z = (idx // (N * N)).astype(np.int32)
    m = r1 - r0
    # neighbours in ascending column order: -N^2, -N, -1, 0, +1, +N, +N^2
    present = np.stack([z > 0, y > 0, x > 0, np.ones(m, bool), x < N - 1, y < N - 1, z < NZ - 1], axis=1)
    del x, y, z
    offs = np.array([-N * N, -N, -1, 0, 1, N, N * N], dtype=np.int64)
    counts = present.sum(axis=1, dtype=np.int64)
    indptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    r, s = np.nonzero(present)
    indices = (r + r0 + offs[s]).astype(np.int32)
    data = np.where(s == 3, 6.0, -1.0).astype(dtype)
    A = sp.csr_matrix((data, indices, indptr.astype(np.int32)), shape=(m, n))
    A.has_sorted_indices = True
    return A


def powerlaw_spd(n: int = 5_000_000, nnz_target: int = 50_000_000, alpha: float = 1.6,
                 max_row: int = 200_000, seed: int = 12345, perm_seed: int = 54321) -> sp.csr_matrix:
    """Irregular SPD matrix with power-law row lengths (BASELINE.json config 5).

    Strict-upper pattern with Pareto(alpha) row degrees rescaled so that the mirrored
    matrix has about `nnz_target` non-zeros including the diagonal, uniform random
    columns, values -U(0,1), diagonal = sum|off-diagonal| + 1 (strictly diagonally
    dominant => SPD), then a symmetric random permutation so long rows are scattered."""
    rng = np.random.default_rng(seed)
    half = (nnz_target - n) // 2
    deg = np.minimum(np.floor(rng.pareto(alpha, n) * 2.0) + 1.0, float(max_row))
    deg = np.maximum(np.rint(deg * (half / deg.sum())), 0).astype(np.int64)
    deg = np.minimum(deg, max_row)
    rows = np.repeat(np.arange(n, dtype=np.int64), deg)
    cols = rng.integers(0, n, rows.size, dtype=np.int64)
    keep = rows != cols
    rows, cols = rows[keep], cols[keep]
    lo, hi = np.minimum(rows, cols), np.maximum(rows, cols)
    vals = -rng.random(lo.size)
    U = sp.coo_matrix((vals, (lo, hi)), shape=(n, n)).tocsr()
    U.sum_duplicates()
    # duplicates were summed: clamp back into (-1, 0) so the spec's value range holds
    np.maximum(U.data, -0.999999, out=U.data)
    A = U + U.T
    diag = np.asarray(abs(A).sum(axis=1)).ravel() + 1.0
    A = A + sp.diags(diag)
    perm = np.random.default_rng(perm_seed).permutation(n)
    A = A.tocsr()[perm][:, perm]
    return _canonical(A)


def row_patterns(A: sp.csr_matrix) -> int:
    """Number of distinct rows of A written as lists of (column - row, value bits) -- what the engine's row-pattern
    dictionary finds on the device.  Host-side restatement for tests and for choosing workloads."""
    A = A.tocsr()
    ip, ix, d = A.indptr, A.indices, A.data
    seen = set()
    for i in range(A.shape[0]):
        lo, hi = ip[i], ip[i + 1]
        seen.add(((ix[lo:hi].astype(np.int64) - i).tobytes(), d[lo:hi].tobytes()))
    return len(seen)


def csr_arrays(A: sp.csr_matrix, dtype: str):
    """(values, rowptr, cols) as contiguous arrays of the ABI types for dtype name `dtype`."""
    np_t = DTYPES[dtype][0]
    vals = A.data if DTYPES[dtype][3] or not np.iscomplexobj(A.data) else A.data.real
    return (np.ascontiguousarray(vals, dtype=np_t),
            np.ascontiguousarray(A.indptr, dtype=np.intc),
            np.ascontiguousarray(A.indices, dtype=np.intc))


def algorithmic_bytes(n: int, nnz: int, k: int, dtype: str):
    """(B_spmv, B_iter) of SURVEY.md 8(d): B_spmv = nnz(v+4) + 4(n+1) + 2knv, B_iter = B_spmv + 9knv."""
    v = DTYPES[dtype][2]
    b_spmv = nnz * (v + 4) + 4 * (n + 1) + 2 * k * n * v
    return b_spmv, b_spmv + 9 * k * n * v


def flops_per_iteration(n: int, nnz: int, k: int, dtype: str) -> int:
    """The report's flop model (Table II): real k(2nnz+10n+2), complex k(8nnz+40n+28)."""
    if DTYPES[dtype][3]:
        return k * (8 * nnz + 40 * n + 28)
    return k * (2 * nnz + 10 * n + 2)
