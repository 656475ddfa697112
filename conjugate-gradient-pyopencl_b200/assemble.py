"""Device-side assembly of the grid operators the reference's drivers build on the host.

`local_rect` (/root/reference/p_helmholtz.py:1342-1542), `helmFE_var` with a constant wave speed
(/root/reference/helmFE_var.py:9-331) and `Poisson` (p_helmholtz.py:1545-1585) are loops over the grid nodes with
one `if` per CLASS of node -- first / interior / last in each direction -- and one fixed list of (neighbour,
coefficient) pairs per class.  A *class table* is that list for the 9 (2-D) or 27 (3-D) classes.  It is a few hundred
bytes; `cgb200_create_grid` (csrc/assemble.cuh) expands it into CSR arrays directly in HBM.

    M = assemble.local_rect(N, k, eps, eta, L, Nhoriz, Nvert)            # -> engine.Matrix on the device
    M = assemble.helmholtz_fe(1024)                                      # class table read off a 5 x 5 instance of the host generator

The coefficients are evaluated HERE, on the host, in the reference's own arithmetic (a handful of Python complex
operations), so that the assembled matrix agrees with the reference's to the last bit; `expand` is the numpy
restatement of the device kernel (the checker of the kernel, and a host generator for the tests).
"""
import ctypes

import numpy as np
import scipy.sparse as sp

try:
    from . import _lib, engine
except ImportError:
    import _lib
    import engine

MAXLEN = 32          # GRID_MAXLEN


class ClassTable:
    """entries[(cz, cy, cx)] = [((dx, dy, dz), value), ...] sorted by (dz, dy, dx); c = 0 first, 1 interior, 2 last."""

    def __init__(self, entries, ndim):
        self.ndim = ndim
        self.entries = {}
        for key, lst in entries.items():
            key = tuple(key)
            key = (0,) * (3 - len(key)) + key
            lst = [((tuple(d) + (0,) * (3 - len(d)))[:3], v) for d, v in lst]
            self.entries[key] = sorted(lst, key=lambda e: (e[0][2], e[0][1], e[0][0]))

    def arrays(self, dtype):
        ln = np.zeros(27, dtype=np.intc)
        dxyz = np.zeros((27, MAXLEN, 3), dtype=np.intc)
        val = np.zeros((27, MAXLEN), dtype=dtype)
        for (cz, cy, cx), lst in self.entries.items():
            c = (cz * 3 + cy) * 3 + cx
            if len(lst) > MAXLEN:
                raise ValueError("a row class has more than 32 entries")
            ln[c] = len(lst)
            for e, (d, v) in enumerate(lst):
                dxyz[c, e] = d
                val[c, e] = v
        return ln, dxyz, val


def _cls(v, size):
    return np.where(v == 0, 0, np.where(v == size - 1, 2, 1))


def expand(table, nx, ny=1, nz=1, dtype=np.complex128):
    """numpy restatement of grid_assemble_kernel: the CSR matrix of `table` on an nx x ny x nz grid (x fastest)."""
    n = nx * ny * nz
    rows = np.arange(n, dtype=np.int64)
    x, y, z = rows % nx, (rows // nx) % ny, rows // (nx * ny)
    cl = (_cls(z, nz) * 3 + _cls(y, ny)) * 3 + _cls(x, nx)
    ln, dxyz, val = table.arrays(dtype)
    counts = ln[cl].astype(np.int64)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    indices = np.empty(indptr[-1], dtype=np.int32)
    data = np.empty(indptr[-1], dtype=dtype)
    for c in np.unique(cl):
        r = rows[cl == c]
        for e in range(ln[c]):
            dx, dy, dz = (int(t) for t in dxyz[c, e])
            indices[indptr[r] + e] = r + dz * nx * ny + dy * nx + dx
            data[indptr[r] + e] = val[c, e]
    A = sp.csr_matrix((data, indices, indptr.astype(np.int32)), shape=(n, n))
    A.has_sorted_indices = True
    return A


def table_from_template(A, dims):
    """Class table of a constant-coefficient grid operator, read off a SMALL instance `A` of it on a grid of `dims`
    (x fastest, every dimension >= 4 so that a first, an interior and a last node exist and offsets decode uniquely)."""
    dims = tuple(int(d) for d in dims)
    nd = len(dims)
    nx, ny, nz = (dims + (1, 1))[:3]
    if any(d < 4 for d in dims):
        raise ValueError("template grid too small")
    A = sp.csr_matrix(A)
    A.sort_indices()
    rep = {0: 0, 1: 1, 2: None}
    entries = {}
    for cz in (range(3) if nd == 3 else (0,)):
        for cy in (range(3) if nd >= 2 else (0,)):
            for cx in range(3):
                pos = [rep[c] if c != 2 else None for c in (cx, cy, cz)]
                X = pos[0] if pos[0] is not None else nx - 1
                Y = (pos[1] if pos[1] is not None else ny - 1) if nd >= 2 else 0
                Z = (pos[2] if pos[2] is not None else nz - 1) if nd == 3 else 0
                row = (Z * ny + Y) * nx + X
                lst = []
                for j in range(A.indptr[row], A.indptr[row + 1]):
                    o = int(A.indices[j]) - row
                    # balanced decode: o = dz*nx*ny + dy*nx + dx with every component in {-1, 0, 1}
                    dz = int(np.rint(o / (nx * ny))) if nd == 3 else 0
                    o2 = o - dz * nx * ny
                    dy = int(np.rint(o2 / nx)) if nd >= 2 else 0
                    dx = o2 - dy * nx
                    if max(abs(dx), abs(dy), abs(dz)) > 1:
                        raise ValueError("not a nearest-neighbour stencil")
                    lst.append(((dx, dy, dz), A.data[j]))
                entries[(cz, cy, cx)] = lst
    return ClassTable(entries, nd)


# ----------------------------------------------------------------------------------------
# the reference's operators as class tables
# ----------------------------------------------------------------------------------------
def local_rect_table(N, k, eps, eta, L=1.0):
    """`local_rect(N, k, eps, eta, L, Nhoriz, Nvert)` (p_helmholtz.py:1342-1542) as a class table: P1 finite elements
    for  -Laplace u - (k^2 + i eps) u = f  with the impedance condition  du/dn - i eta u = 0, mesh width h = L/(N-1),
    node (j, m) = row m*Nhoriz + j.  Every coefficient is evaluated with the expression (and therefore the rounding)
    of the reference: corner / edge / interior diagonal (:1397-1441), axis neighbours on the boundary and inside
    (:1443-1524), the two diagonal neighbours of the triangulation (:1449-1453, :1470-1474, :1482-1485, :1503-1507,
    :1527-1540)."""
    # (plain Python floats / complex, as the drivers pass them: numpy's complex scalar division rounds differently from
    #  Python's in the last bit, and the reference's own result depends on which of the two it is called with)
    N, k, eps, eta, L = int(N), float(k), float(eps), float(eta), float(L)
    h = L * 1.0 / (N - 1.)
    h2 = h ** 2
    k2 = k ** 2
    z = k2 + 1j * eps
    d_corner_a = 1. - z * h2 / 6. - 1j * eta * 2 * h / 3.          # SW and NE corners: two triangles meet
    d_corner_b = 1. - z * h2 / 12. - 1j * eta * 2 * h / 3.         # SE and NW corners: one triangle
    d_edge = 2. - z * h2 / 4. - 2. * 1j * eta * h / 3.
    d_int = 4. - z * h2 / 2.
    ax_bnd = -1. / 2. - z * h2 / 24. - 1j * eta * h / 6.           # neighbour ALONG a boundary edge
    ax_int = -1. - z * h2 / 12.
    dg = -(z * h2) / 12.                                           # NE / SW neighbour through one triangle
    entries = {}
    for cy in range(3):
        for cx in range(3):
            lst = []
            on_x, on_y = cx != 1, cy != 1                          # on the left/right, bottom/top boundary
            if on_x and on_y:
                diag = d_corner_a if (cx, cy) in ((0, 0), (2, 2)) else d_corner_b
            elif on_x or on_y:
                diag = d_edge
            else:
                diag = d_int
            lst.append(((0, 0), diag))
            if cx != 2:
                lst.append(((1, 0), ax_bnd if on_y else ax_int))   # east
            if cx != 0:
                lst.append(((-1, 0), ax_bnd if on_y else ax_int))  # west
            if cy != 2:
                lst.append(((0, 1), ax_bnd if on_x else ax_int))   # north
            if cy != 0:
                lst.append(((0, -1), ax_bnd if on_x else ax_int))  # south
            if cx != 2 and cy != 2:
                lst.append(((1, 1), dg))                           # north-east
            if cx != 0 and cy != 0:
                lst.append(((-1, -1), dg))                         # south-west
            entries[(cy, cx)] = lst
    return ClassTable(entries, 2)


def poisson2d_table():
    """`Poisson(N)` (p_helmholtz.py:1545-1585): diag 4, the four axis neighbours -1."""
    entries = {}
    for cy in range(3):
        for cx in range(3):
            lst = [((0, 0), 4.0)]
            lst += [((1, 0), -1.0)] * (cx != 2) + [((-1, 0), -1.0)] * (cx != 0)
            lst += [((0, 1), -1.0)] * (cy != 2) + [((0, -1), -1.0)] * (cy != 0)
            entries[(cy, cx)] = lst
    return ClassTable(entries, 2)


def laplace3d_table():
    """The 3-D extension BASELINE.json configs 3 and 4 name: diag 6, the six axis neighbours -1."""
    entries = {}
    for cz in range(3):
        for cy in range(3):
            for cx in range(3):
                lst = [((0, 0, 0), 6.0)]
                for a, c in enumerate((cx, cy, cz)):
                    for s in (1, -1):
                        if (s == 1 and c != 2) or (s == -1 and c != 0):
                            d = [0, 0, 0]
                            d[a] = s
                            lst.append((tuple(d), -1.0))
                entries[(cz, cy, cx)] = lst
    return ClassTable(entries, 3)


# ----------------------------------------------------------------------------------------
# device matrices
# ----------------------------------------------------------------------------------------
class GridMatrix(engine.Matrix):
    """An engine.Matrix whose CSR arrays were generated on the device (`cgb200_create_grid`)."""

    def __init__(self, table, nx, ny=1, nz=1, dtype=np.complex128, device=0):
        self.dtype = np.dtype(dtype)
        ln, dxyz, val = table.arrays(self.dtype)
        self.code = _lib.DTYPE_CODE[self.dtype]
        self.device = int(device)
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().cgb200_create_grid(ctypes.byref(h), self.code, self.device, int(nx), int(ny), int(nz),
                                                 _lib.ptr(ln), _lib.ptr(dxyz), _lib.ptr(val)))
        self._h = h
        info = self.info()
        self.n, self.nnz = int(info["n"]), int(info["nnz"])

    def to_scipy(self):
        """The assembled arrays, copied back (`cgb200_read_matrix`)."""
        vals = np.empty(self.nnz, dtype=self.dtype)
        ptr = np.empty(self.n + 1, dtype=np.intc)
        cols = np.empty(self.nnz, dtype=np.intc)
        _lib.check(_lib.lib().cgb200_read_matrix(self._h, _lib.ptr(vals), _lib.ptr(ptr), _lib.ptr(cols)))
        A = sp.csr_matrix((vals, cols, ptr), shape=(self.n, self.n))
        A.has_sorted_indices = True
        return A


def local_rect(N, k, eps, eta, L, Nhoriz, Nvert, dtype=np.complex128, device=0):
    """The subdomain operator of as_prec (p_h-PY_C-CL.py:1881,1906 call local_rect(N, k, eps, eta=k, L, Nhoriz, Nvert)),
    assembled on the device."""
    return GridMatrix(local_rect_table(N, k, eps, eta, L), Nhoriz, Nvert, 1, dtype=dtype, device=device)


def poisson2d(N, dtype=np.float64, device=0):
    return GridMatrix(poisson2d_table(), N, N, 1, dtype=dtype, device=device)


def laplace3d(N, nz=None, dtype=np.float64, device=0):
    return GridMatrix(laplace3d_table(), N, N, N if nz is None else nz, dtype=dtype, device=device)


def helmholtz_fe(N, omega=12.0, rho=0.15, dtype=np.complex128, device=0):
    """`helmFE_var(N, omega, C=ones, rho, N, N)` (helmFE_var.py:9-331) for a CONSTANT wave speed, assembled on the
    device.  The class table is read off the host generator's own 5 x 5 instance with the mesh width of the N x N grid
    (problems.helmholtz_fe(5, ..., h=1/(N-1)) -- checked bit for bit against the reference function), so the
    coefficients carry the reference's rounding.  (The reference special-cases the rounding of the top-right corner,
    helmFE_var.py:104: that node is a class of its own here as well.)"""
    try:
        from . import problems
    except ImportError:
        import problems
    T = problems.helmholtz_fe(5, omega=omega, rho=rho, h=1.0 / (N - 1.0))
    return GridMatrix(table_from_template(T, (5, 5)), N, N, 1, dtype=dtype, device=device)
