"""Row-block sharded CG on N GPUs (torchrun): parity with the single-GPU engine and the oracle,
then timing.  python -m torch.distributed.run --nproc-per-node N tools/shard_check.py [--parity-only] [N3d] [iters]

The parity half is what tests/test_gpu_cg2.py::test_row_block_sharded_cg_parity_over_nvlink runs on 2/4/8 GPUs."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, torch.distributed as dist
import cg_b200
from cg_b200 import sharded
import cg_b200.problems as P

import faulthandler
faulthandler.dump_traceback_later(150, exit=True)
PARITY_ONLY = "--parity-only" in sys.argv
ARGS = [a for a in sys.argv[1:] if not a.startswith("--")]


def log(*a):
    print(f"[rank {os.environ.get('RANK', 0)} +{time.time() - T0:.1f}s]", *a, file=sys.stderr, flush=True)


T0 = time.time()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
out = {}
# ---- parity on small systems (every rank builds the global matrix)
import cpu_ref
for name, A, b, tolv in (("lap3d24", P.laplace3d(24), np.ones(24 ** 3), 1e-10),
                         ("helm64_c128", P.helmholtz_fe(64), P.rhs_a(64, 12.0), 1e-10),
                         ("powerlaw", P.powerlaw_spd(n=20000, nnz_target=300000, max_row=3000), None, 1e-10)):
    if b is None:
        b = A @ np.ones(A.shape[0])
    by = "nnz" if name == "powerlaw" else "rows"
    bounds = sharded.split_rows(A.indptr, world, by=by)
    log(name, "planning")
    plan = sharded.plan_row_block(A.indptr, A.indices, A.data, bounds, rank)
    rb, re = bounds[rank], bounds[rank + 1]
    log(name, "plan done; creating shard", plan.n_owned, plan.n_halo)
    M = sharded.ShardedMatrix(plan, device=lr)
    log(name, "shard created; plain-launch solve", "p2p" if M.p2p else "nccl")
    M.set_option("use_graph", 0)
    x, info = M.solve(b[rb:re].astype(A.dtype), max_iterations=60)
    cg2 = M.get_option("cg2_ok")
    if cg2 and M.get_option("march_ok"):
        # the plane-marching dir_spmv (halo planes as the pieces below / above the owned planes) vs the window kernel
        M.set_option("march", 2)
        xm, _ = M.solve(b[rb:re].astype(A.dtype), max_iterations=60)
        M.set_option("march", 0)
        xw, _ = M.solve(b[rb:re].astype(A.dtype), max_iterations=60)
        M.set_option("march", 1)
        em = float(np.linalg.norm(xm - xw) / np.linalg.norm(xw))
        log(name, "plane-marching vs window dir_spmv:", em)
        assert em < 1e-10, em
    if cg2:
        # the two-kernel iteration (halo stores from inside the producing kernels) vs the three-kernel one
        # (halo_push_kernel on a side stream + arrival flags): same FMAs, other association of the dot sums
        M.set_option("cg2", 0)
        x3, _ = M.solve(b[rb:re].astype(A.dtype), max_iterations=60)
        M.set_option("cg2", 1)
        e3 = float(np.linalg.norm(x - x3) / np.linalg.norm(x3))
        log(name, "two-kernel vs three-kernel iteration:", e3)
        assert e3 < 1e-10, e3
    if world > 1:
        M.enable_peer_memory(False)
        xn, _ = M.solve(b[rb:re].astype(A.dtype), max_iterations=60)
        M.enable_peer_memory(True)
        en = float(np.linalg.norm(x - xn) / np.linalg.norm(xn))
        log(name, "peer-memory vs NCCL collectives:", en)
        assert en < 1e-12, en
    M.set_option("pattern", 0)                       # CSR kernels instead of the row-pattern dictionary
    xc, _ = M.solve(b[rb:re].astype(A.dtype), max_iterations=60)
    M.set_option("pattern", 1)
    ec = float(np.linalg.norm(x - xc) / np.linalg.norm(xc))
    log(name, "pattern dictionary vs CSR kernels:", ec)
    assert ec < 1e-11, ec
    log(name, "plain solve done; graph solve")
    M.set_option("use_graph", 1)
    xg, _ = M.solve(b[rb:re].astype(A.dtype), max_iterations=60)
    assert np.array_equal(x, xg)
    log(name, "graph solve done")
    ref, _, _ = cpu_ref.cg(A.data, A.indptr, A.indices, b.astype(A.dtype), iters=60)
    err = float(np.linalg.norm(x - ref[rb:re]) / np.linalg.norm(ref[rb:re]))
    log(name, "oracle done; tolerance solve, plain launches")
    M.set_option("use_graph", 0)
    x2p, info2p = M.solve(b[rb:re].astype(A.dtype), max_iterations=5000, tol=1e-9)
    log(name, "tolerance solve, graphs", info2p["iterations"])
    M.set_option("use_graph", 1)
    x2, info2 = M.solve(b[rb:re].astype(A.dtype), max_iterations=5000, tol=1e-9)
    log(name, "tolerance solves done", info2["iterations"])
    # (this pair once differed by an iteration: the second solve's first SpMV ran ahead of its halo)
    log(name, "plain vs graph tolerance solve: iterations", info2p["iterations"], info2["iterations"],
        "x difference", float(np.linalg.norm(x2 - x2p) / np.linalg.norm(x2p)))
    assert np.array_equal(x2, x2p) and info2["iterations"] == info2p["iterations"]
    _, its_ref, _ = cpu_ref.cg(A.data, A.indptr, A.indices, b.astype(A.dtype), iters=5000, tol=1e-9)
    out[name] = dict(err60=err, iters=info2["iterations"], iters_oracle=int(its_ref[0]), n_halo=plan.n_halo, info=M.info(),
                     two_kernel=int(cg2), march_ok=int(M.get_option("march_ok")))
    assert err < tolv, (name, err)
    assert abs(info2["iterations"] - int(its_ref[0])) <= 1, (name, info2, its_ref)
    M.close()
if PARITY_ONLY:
    if rank == 0:
        print(json.dumps({"world": world, **out}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0)
# ---- timing: 3-D Laplacian N^3, each rank builds only its slab
N3 = int(ARGS[0]) if len(ARGS) > 0 else 300
iters = int(ARGS[1]) if len(ARGS) > 1 else 256
n = N3 ** 3
planes = [(N3 * p) // world for p in range(world + 1)]
bounds = np.array([pl * N3 * N3 for pl in planes], dtype=np.int64)
rb, re = int(bounds[rank]), int(bounds[rank + 1])
t0 = time.time()
Al = P.laplace3d(N3, rows=(rb, re))
plan = sharded.plan_row_block(Al.indptr, Al.indices, Al.data, bounds, rank)
tgen = time.time() - t0
log("timing: shard create")
M = sharded.ShardedMatrix(plan, device=lr)
faulthandler.cancel_dump_traceback_later()
faulthandler.dump_traceback_later(200, exit=True)
bt = torch.ones(re - rb, dtype=torch.float64, device="cuda")
xt = torch.zeros_like(bt)
for use_graph, p2p in ((1, 1), (1, 0), (0, 1)):
    if world > 1:
        M.enable_peer_memory(bool(p2p))
    elif not p2p:
        continue
    M.set_option("use_graph", use_graph)
    for _ in range(2):
        xt.zero_(); M.solve(bt, xt, max_iterations=iters)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        xt.zero_(); _, info = M.solve(bt, xt, max_iterations=iters); ts.append(info["timing_ms"]["iterations"])
    t = torch.tensor([min(ts)], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[f"lap{N3}_graph{use_graph}_p2p{p2p}"] = dict(ms_per_iter=float(t.item()) / iters, its_per_s=iters / float(t.item()) * 1e3,
                                            n_owned=plan.n_owned, n_halo=plan.n_halo, gen_s=tgen)
out["local_kernels_us"] = {nm: round(1e3 * M.time_kernel(nm, reps=100), 2)
                           for nm in ("spmv_dot", "update_xr", "update_d") + (("dir_spmv", "update_r") if M.get_option("cg2_ok") else ())}
M.close()
if rank == 0:
    print(json.dumps({"world": world, **out}))
if world > 1:
    dist.destroy_process_group()
