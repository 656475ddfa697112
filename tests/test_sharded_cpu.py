"""Host-side logic of the multi-GPU modes on CPU: world_size 2 and 3 over gloo.

  - rhs-split reproduces the reference's distribution (p_h-PY_C-CL-multi-GPU.py:2123-2140);
  - the row-block plan (local renumbering, halo lists, who-sends-what) is exercised by running the
    exact communication pattern of the CUDA path in numpy: halo exchange + two all-reduces per
    iteration must reproduce the global CG oracle.
"""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _matrix(kind):
    import cg_b200.problems as P
    if kind == "helm":
        return P.helmholtz_fe(20), P.rhs_a(20, 12.0)
    if kind == "lap3d":
        A = P.laplace3d(9)
        return A, np.ones(A.shape[0])
    # irregular SPD: power-law rows, halo ~ everything
    A = P.powerlaw_spd(n=600, nnz_target=6000, max_row=200)
    return A, A @ np.ones(A.shape[0])


def _worker(rank, world, port, kind, by, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cg_b200 import sharded
        import sharded_numpy
        import np_cg
        A, b = _matrix(kind)
        n = A.shape[0]
        bounds = sharded.split_rows(A.indptr, world, by=by)
        rb, re = bounds[rank], bounds[rank + 1]
        plan = sharded.plan_row_block(A.indptr, A.indices, A.data, bounds, rank)
        # a rank that only holds its own slab builds the same plan
        sl = slice(A.indptr[rb], A.indptr[re])
        plan2 = sharded.plan_row_block(A.indptr[rb:re + 1] - A.indptr[rb], A.indices[sl], A.data[sl], bounds, rank)
        for f in ("indptr", "cols_local", "send_idx", "send_counts", "recv_counts"):
            assert np.array_equal(getattr(plan, f), getattr(plan2, f)), f
        assert plan.n_owned == re - rb and plan.recv_counts[rank] == 0 and plan.send_counts[rank] == 0
        touches = np.array([(plan.cols_local[plan.indptr[i]:plan.indptr[i + 1]] >= plan.n_owned).any()
                            for i in range(plan.n_owned)])
        assert np.array_equal(plan.row_boundary.astype(bool), touches)
        assert plan.recv_counts.sum() == plan.n_halo and plan.send_counts.sum() == plan.send_idx.size
        # halo exchange delivers exactly the referenced remote entries
        v = (np.arange(n) * 1.5 + 0.25).astype(A.dtype)
        loc = sharded_numpy.halo_exchange_numpy(plan, v[rb:re], dist)
        assert np.array_equal(loc[:plan.n_owned], v[rb:re])
        assert np.array_equal(loc[plan.n_owned:], v[plan.halo_globals])
        # local SpMV == rows of the global one
        Al = sp.csr_matrix((plan.data, plan.cols_local, plan.indptr), shape=(plan.n_owned, plan.n_owned + plan.n_halo))
        assert np.allclose(Al @ loc, (A @ v)[rb:re], rtol=1e-13, atol=1e-13)
        # sharded CG == the global oracle
        its = 25
        x = sharded_numpy.reference_sharded_cg(plan, b[rb:re].astype(A.dtype), np.zeros(re - rb, A.dtype), its, dist)
        ref = np_cg.cg(A, b.astype(A.dtype), x=np.zeros(n, A.dtype), maxit=its)
        err = np.linalg.norm(x - ref[rb:re]) / np.linalg.norm(ref[rb:re])
        assert err < 1e-11, err
        open(os.path.join(out_dir, f"ok{rank}"), "w").write(f"{plan.n_owned} {plan.n_halo} {err:.2e}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kind,by", [(2, "helm", "rows"), (3, "lap3d", "rows"), (2, "powerlaw", "nnz"),
                                           (3, "helm", "nnz")])
def test_row_block_plan_and_exchange_over_gloo(tmp_path, world, kind, by):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), kind, by, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == [f"ok{r}" for r in range(world)]


def test_split_rows():
    from cg_b200 import sharded
    import cg_b200.problems as P
    A = P.laplace3d(10)
    b = sharded.split_rows(A.indptr, 4)
    assert list(b) == [0, 250, 500, 750, 1000]
    b = sharded.split_rows(A.indptr, 8, align=100)          # whole z-planes (300 planes / 8 ranks)
    assert all(x % 100 == 0 for x in b) and b[-1] == 1000 and len(b) == 9
    A = P.powerlaw_spd(n=2000, nnz_target=30000, max_row=500)
    b = sharded.split_rows(A.indptr, 4, by="nnz")
    per = np.diff(A.indptr[b])
    assert per.max() < 1.25 * per.mean()                     # balanced by non-zeros, not by rows
    with pytest.raises(ValueError):
        sharded.split_rows(np.arange(3), 5)


def test_rhs_split_matches_reference_distribution():
    from cg_b200 import sharded
    # p_h-PY_C-CL-multi-GPU.py:2125-2134: n // g each, the first n % g devices one more, contiguous
    assert sharded.split_rhs(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sharded.split_rhs(16, 8) == [(2 * i, 2 * i + 2) for i in range(8)]
    assert sharded.split_rhs(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    for n in range(1, 40):
        for g in range(1, 9):
            parts = sharded.split_rhs(n, g)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [e - s for s, e in parts]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@pytest.mark.reference
def test_rhs_split_against_lifted_reference_function():
    import ast
    from cg_b200 import sharded
    src = open("/root/reference/p_h-PY_C-CL-multi-GPU.py").read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "distribute_workloads_on_devices")

    class FakePcl:
        @staticmethod
        def initialize_cl_environment_with_device(d):
            return ("ctx", d), ("queue", d)

        @staticmethod
        def load_and_build_kernels(ctx, k):
            return {"n_rhs": k}
    env = {"pcl": FakePcl}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref", "exec"), env)
    for n, g in ((10, 4), (64, 8), (5, 8), (9, 2)):
        ref = env["distribute_workloads_on_devices"](list(range(g)), n)
        assert [(ref[d][0], ref[d][1]) for d in range(g)] == sharded.split_rhs(n, g)
        assert all(ref[d][4]["n_rhs"] == ref[d][1] - ref[d][0] for d in range(g))
