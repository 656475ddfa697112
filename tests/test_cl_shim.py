"""Host logic of the `cl` drop-in module (no GPU needed): both call forms of the reference
drivers parse to the same C call; devices / contexts / kernels are the expected tokens."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def pcl(monkeypatch):
    # the reference drivers do `import cl as pcl` with the module directory on sys.path
    monkeypatch.syspath_prepend(os.path.join(ROOT, "conjugate-gradient-pyopencl_b200"))
    sys.modules.pop("cl", None)
    import cl
    yield cl
    sys.modules.pop("cl", None)


def test_module_surface_matches_reference(pcl):
    for name in ("initialize_cl_environment", "initialize_cl_environment_with_device", "get_gpu_devices",
                 "load_and_build_kernels", "create_kernels", "CG", "conjugate_gradient_multi_gpu",
                 "IS_COMPLEX", "WAVE_SIZE", "LOCAL_SIZE"):
        assert hasattr(pcl, name), name
    assert pcl.LOCAL_SIZE == 256 and pcl.WAVE_SIZE == 32
    k = pcl.load_and_build_kernels(None, 4)
    assert set(k) == {"axpy", "aypx", "spmv", "sub", "vdot"}        # cl.py:36-42
    assert set(pcl.create_kernels(1)) == set(k)


def test_both_call_forms_reach_the_same_c_call(pcl, monkeypatch):
    calls = []
    monkeypatch.setattr(pcl, "_solve", lambda dev, *a: calls.append((dev, a)) or a[6])
    ctx, queue = pcl.initialize_cl_environment_with_device(pcl.Device(3))
    kern = pcl.load_and_build_kernels(ctx, 2)
    a = np.ones(3, np.csingle); b = np.ones(4, np.csingle); p = np.zeros(3, np.intc); c = np.zeros(3, np.intc)
    x = np.zeros(4, np.csingle)
    assert pcl.CG(ctx, queue, kern, 2, 3, a, b, p, c, x, 2, 7) is x                 # p_h-PY_C-CL.py:1933
    assert pcl.CG(2, 3, a, b, p, c, x, 2, 7) is x                                   # p_helmholtz.py:1839
    assert pcl.CG(2, 3, a, b, p, c, x, 2, 7, pcl.Device(1)) is x                    # commented form, :1934
    assert pcl.conjugate_gradient_multi_gpu(ctx, queue, kern, 2, 3, a, b, p, c, x, 2, 7, pcl.Device(5)) is x
    assert [d for d, _ in calls] == [3, 0, 1, 5]
    assert all(args[0] == 2 and args[1] == 3 and args[7] == 2 and args[8] == 7 for _, args in calls)
    with pytest.raises(TypeError):
        pcl.CG(1, 2, 3)


def test_input_validation(pcl):
    a = np.ones(3, np.csingle); p = np.array([0, 1, 3], np.intc); c = np.zeros(3, np.intc)
    with pytest.raises(TypeError):       # x of the wrong dtype cannot be filled in place
        pcl.CG(2, 3, a, np.ones(2, np.csingle), p, c, np.zeros(2, np.complex128), 1, 1)
    with pytest.raises(ValueError):      # b shorter than n_rhs * size
        pcl.CG(2, 3, a, np.ones(2, np.csingle), p, c, np.zeros(4, np.csingle), 2, 1)
    with pytest.raises(TypeError):
        pcl.CG(2, 3, a.astype(np.int32), np.ones(2), p, c, np.zeros(2), 1, 1)


def test_oclcgex_cli_arguments(tmp_path, capsys):
    """The example executable's four arguments and error paths (main.c:15-24), without a GPU."""
    from cg_b200 import oclcgex
    assert oclcgex.main(["only", "three", "args"]) == 1
    assert "Usage: ./CG <input matrix file>" in capsys.readouterr().err
    assert oclcgex.main([str(tmp_path / "missing.mtx"), "1", "1", "10"]) == 1
    assert "Could not read matrix" in capsys.readouterr().out
    # a symmetric Matrix Market file is expanded to full storage, as main.c:25 does
    import scipy.io, scipy.sparse as sp
    A = sp.csr_matrix(np.array([[4.0, -1, 0], [-1, 4, -1], [0, -1, 4]]))
    scipy.io.mmwrite(str(tmp_path / "sym.mtx"), sp.tril(A), symmetry="symmetric")
    B = oclcgex.load_csr(str(tmp_path / "sym.mtx"))
    assert abs(B - A).max() == 0 and B.has_sorted_indices


def _native_oclcgex():
    import cg_b200.build as B
    B.build()
    return B.EXE


def test_native_oclcgex_arguments_and_matrix_market_errors(tmp_path):
    """build/oclcgex (csrc/oclcgex.c), the compiled twin of main.c: usage (main.c:15-18), unreadable file
    (main.c:21-24), a complex file with <is complex> = 0 -- all before any device work."""
    import subprocess
    import scipy.io, scipy.sparse as sp
    exe = _native_oclcgex()
    r = subprocess.run([exe, "a", "b"], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage: ./CG <input matrix file>" in r.stderr
    r = subprocess.run([exe, str(tmp_path / "missing.mtx"), "1", "1", "10"], capture_output=True, text=True)
    assert r.returncode == 1 and "Could not read matrix" in r.stdout
    (tmp_path / "junk.mtx").write_text("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n")
    r = subprocess.run([exe, str(tmp_path / "junk.mtx"), "1", "0", "10"], capture_output=True, text=True)
    assert r.returncode == 1 and "Could not read matrix" in r.stdout          # dense `array` files are not sparse matrices
    A = sp.csr_matrix(np.array([[4.0 + 1j, -1], [-1, 4.0]]))
    scipy.io.mmwrite(str(tmp_path / "c.mtx"), A)
    r = subprocess.run([exe, str(tmp_path / "c.mtx"), "1", "0", "10"], capture_output=True, text=True)
    assert r.returncode == 1 and "matrix is complex" in r.stdout


@pytest.mark.parametrize("field,symmetry", [("real", "general"), ("real", "symmetric"), ("complex", "symmetric"),
                                            ("complex", "hermitian"), ("real", "skew-symmetric"), ("pattern", "symmetric"),
                                            ("integer", "general")])
def test_native_matrix_market_reader(tmp_path, field, symmetry):
    """csrc/oclcgex.c reads the coordinate format itself (the reference uses BeBOP SMC, main.c:20-27): every
    field / symmetry combination must expand to the same full-storage CSR matrix scipy's reader gives, with
    sorted rows and summed duplicates -- what cg() then receives."""
    import subprocess
    import scipy.io, scipy.sparse as sp
    exe = _native_oclcgex()
    rng = np.random.default_rng(hash((field, symmetry)) % 2**32)
    n = 9
    L = sp.random(n, n, density=0.3, random_state=3, format="coo")
    L = sp.tril(L, k=-1 if symmetry == "skew-symmetric" else 0).tocoo()
    vals = rng.integers(1, 9, L.nnz).astype(float)
    if field == "complex":
        vals = vals + 1j * rng.integers(1, 9, L.nnz)
        if symmetry == "hermitian":
            vals = np.where(L.row == L.col, vals.real, vals)
    path = tmp_path / "m.mtx"
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate {field} {symmetry}\n% a comment\n\n{n} {n} {L.nnz + (1 if symmetry == 'general' else 0)}\n")
        for r, c, v in zip(L.row, L.col, vals):
            f.write(f"{r + 1} {c + 1}" + ("" if field == "pattern" else (f" {v.real:g} {v.imag:g}" if field == "complex" else
                                                                    (f" {int(v.real)}" if field == "integer" else f" {v.real:.17g}"))) + "\n")
        if symmetry == "general":          # a duplicate entry: summed
            f.write(f"{L.row[0] + 1} {L.col[0] + 1}" + (" 2" if field == "integer" else " 2.5") + "\n")
    out = subprocess.run([exe, str(path), "--dump-csr"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.split("\n")
    hn, hnnz, hc = (int(v) for v in lines[0].split())
    ent = np.array([[float(v) for v in l.split()] for l in lines[1:] if l.strip()]).reshape(-1, 4)
    got = sp.csr_matrix((ent[:, 2] + 1j * ent[:, 3], (ent[:, 0].astype(int), ent[:, 1].astype(int))), shape=(n, n))
    ref = sp.csr_matrix(scipy.io.mmread(str(path)))
    ref.sum_duplicates()
    assert hn == n and hnnz == ref.nnz == len(ent) and hc == (field == "complex")
    assert abs(got - ref).max() == 0
    rows, cols = ent[:, 0].astype(int), ent[:, 1].astype(int)
    assert np.all(np.diff(rows) >= 0) and np.all((np.diff(cols) > 0) | (np.diff(rows) > 0))      # CSR order, no duplicates
