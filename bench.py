#!/usr/bin/env python3
"""bench.py -- CG iterations/s of the B200 engine on the BASELINE.json configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--dtype c128]
    python bench.py --impl reference ...        # the CPU oracle port on the host cores

A *step* is one call of the hot path as the reference's caller makes it: one
`cg(size, nnz, A, b, ptr, cols, x, k, nIterations=256, isComplex)` (CGMaxIT = 256 is the
driver default, p_h-PY_C-CL.py:3553), i.e. 256 CG iterations on one batch of k right-hand sides.

  value   iterations/s with the matrix, b and x already resident in HBM (device pointers through
          cgb200_solve), CUDA events on the launching stream, max over ranks.
  e2e     the same metric through the exported `cg`/`cgd` symbol with pinned HOST buffers: every step uploads b and
          x0 and reads x back; the matrix of the previous call stays resident (content-hashed every call) -- the same
          kind of number as the sharded e2e at N > 1; `e2e.matrix_uploaded` has the matrix uploaded inside every call.
  roofline  the kernel that takes the most time, timed alone with CUDA events on the engine's stream; algorithmic
          bytes per SURVEY.md 8(d) (B_spmv = nnz(v+4) + 4(n+1) + 2knv, + 9knv for the vector passes) against
          MEASURED_PEAKS.json, and the bytes the kernel really moves (moved_*) where a format compresses them.
  cpu_baseline  the oracle port (oracle/cpu_ref.c, OpenMP, all host cores) on a bounded sample.

Default workload: C4, the 3-D 7-point Laplacian 300^3 (27 M unknowns, f64) -- the configuration
BASELINE.json's metric ("... at 1/2/4/8 B200") and scaling target are quoted on; it fits one GPU, so the
same system is measured at every N (strong scaling).  At N = 1 the line also carries the C2 figures
(complex-symmetric Helmholtz FE 1024^2, c128) under "also".

N > 1 (torchrun, one rank per GPU): the CSR matrix is row-block partitioned (whole z-planes per rank);
every iteration moves only the halo of d peer to peer over NVLink and all-reduces the two dot scalars
(--mode rhs-split: the reference's own multi-GPU mode, RHS columns split, no collective,
p_h-PY_C-CL-multi-GPU.py:2123-2181).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ITERS_PER_STEP = 256
WORKLOADS = {
    "c1": dict(desc="C1 2-D 5-point Poisson 256x256", dtype="f64", k=1),
    "c2": dict(desc="C2 complex-symmetric Helmholtz FE 1024x1024 (helmFE_var, omega=12, rho=0.15)", dtype="c128", k=1),
    "c3": dict(desc="C3 3-D 7-point Laplacian 128^3, 32 right-hand sides", dtype="f64", k=32),
    "c4": dict(desc="C4 3-D 7-point Laplacian 300^3", dtype="f64", k=1),
    "c5": dict(desc="C5 power-law SPD, 5M rows, 50M nnz", dtype="f64", k=1),
    # not a BASELINE config: what ONE of 8 GPUs holds of C4 (38 of the 300 z-planes), to tune the shard-sized
    # kernels on a single GPU
    "c4slab8": dict(desc="one eighth of C4: 3-D 7-point Laplacian 300x300x38", dtype="f64", k=1),
    # ... and one quarter of it (76 planes): on TWO GPUs every rank holds exactly what a rank of the 8-GPU run of C4
    # holds, so the shard-sized kernels and the exchange can be tuned at a quarter of the GPU time
    "c4slab4": dict(desc="one quarter of C4: 3-D 7-point Laplacian 300x300x76", dtype="f64", k=1),
}


def make_problem(name, dtype, k):
    """make_problem_uncached, or its result kept under $CGB200_PROBLEM_CACHE (a directory) between the
    runs of one measuring session -- generating C4 / C5 takes longer than benching them."""
    import scipy.sparse as sp
    cache = os.environ.get("CGB200_PROBLEM_CACHE")
    if not cache:
        return make_problem_uncached(name, dtype, k)
    path = os.path.join(cache, f"{name}_{dtype}_{k}.npz")
    if os.path.exists(path):
        z = np.load(path)
        A = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
        A.has_sorted_indices = True
        return A, z["B"]
    A, B = make_problem_uncached(name, dtype, k)
    os.makedirs(cache, exist_ok=True)
    np.savez(path, data=A.data, indices=A.indices, indptr=A.indptr, shape=np.array(A.shape), B=B)
    return A, B


def make_problem_uncached(name, dtype, k):
    """(A scipy CSR, B flat [k][n]) of BASELINE.json config `name` (SURVEY.md 8(d) inputs)."""
    import cg_b200.problems as P
    np_t = P.DTYPES[dtype][0]
    cplx = P.DTYPES[dtype][3]
    if name == "c1":
        A = P.poisson2d(256)
        b = np.ones(A.shape[0])
    elif name == "c2":
        A = P.helmholtz_fe(1024)
        b = P.rhs_a(1024, 12.0)
    elif name == "c3":
        A = P.laplace3d(128)
        b = None
    elif name == "c4":
        A = P.laplace3d(300)
        b = np.ones(A.shape[0])
    elif name in ("c4slab8", "c4slab4"):
        A = P.laplace3d(300, nz=38 if name == "c4slab8" else 76)
        b = np.ones(A.shape[0])
    elif name == "c5":
        A = P.powerlaw_spd()
        # (SURVEY.md 8(d) suggests b = A.ones, but every row of this matrix sums to 1, so ones is an
        #  eigenvector and CG converges in one iteration; a random solution keeps the 256 iterations busy)
        b = A @ np.random.default_rng(7).uniform(-1.0, 1.0, A.shape[0])
    else:
        raise SystemExit(f"unknown workload {name}")
    n = A.shape[0]
    if not cplx and np.iscomplexobj(A.data):
        raise SystemExit(f"workload {name} is complex; pick --dtype c64 or c128")
    if b is None or k > 1:
        cols = []
        for r in range(k):
            if b is not None and r == 0:
                cols.append(b)
            else:
                v = np.random.default_rng(1000 + r).uniform(-1.0, 1.0, n)
                cols.append(v + 0j if cplx else v)
        B = np.concatenate(cols)
    else:
        B = b
    A = A.astype(np_t)
    return A, np.ascontiguousarray(B, dtype=np_t)


def summarise_clock_rows(rows, window=None, post=None):
    """nvidia-smi rows ("timestamp, index, clocks.sm, clocks.max.sm, power.draw, reasons.active, hw_slowdown, hw_thermal_slowdown,
    sw_thermal_slowdown, sw_power_cap") -> the `clocks` object of the bench line.  `window` = (t0, t1) wall-clock seconds of the
    timed region: only samples inside it count.  When it holds none (a region shorter than nvidia-smi's ~100 ms period: the
    sharded runs at 4 and 8 GPUs), the samples inside `post` -- extra untimed steps under the same load, run right after
    -- are used instead and the object says so."""
    import datetime
    names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
    parsed = []
    for r in rows:
        if len(r) < 10:
            continue
        try:
            ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
            parsed.append((ts, float(r[2]), float(r[3]), float(r[4]), [nm for nm, v in zip(names, r[6:10]) if v.strip().lower().startswith("active")]))
        except ValueError:
            continue

    def inside(w):
        return [p for p in parsed if w is None or (w[0] - 0.02 <= p[0] <= w[1] + 0.02)]
    out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    use, where = inside(window), "timed region"
    if not use and post is not None:
        use, where = inside(post), "post-roll: untimed steps under the same load right after the timed region (shorter than the sampling period)"
    if not use and parsed and window is not None:
        # nvidia-smi's clock and the host's did not line up (time zone of the tool's output): the sampler is stopped right
        # after the measured steps, so the LAST samples are the ones taken under load
        span = (post[1] - post[0]) if post is not None else (window[1] - window[0])
        use, where = parsed[-max(1, min(len(parsed), int(span / 0.1))):], "last samples before the sampler was stopped (timestamps did not match the host clock)"
    if use:
        out.update(sm_mhz=statistics.median(p[1] for p in use), sm_max_mhz=max(p[2] for p in use),
                   reasons=sorted(set(x for p in use for x in p[4])), samples=len(use), power_w_max=max(p[3] for p in use),
                   sampled_in=where)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  Started early (nvidia-smi
    needs ~0.5 s to come up); begin() / end() bracket the timed region, post_begin() / post_end() an optional post-roll."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t0 = self.t1 = self.p0 = self.p1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def post_begin(self):
        self.p0 = time.time()

    def post_end(self):
        self.p1 = time.time()

    def timed_seconds(self):
        return (self.t1 - self.t0) if (self.t0 is not None and self.t1 is not None) else None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        window = (self.t0, self.t1) if self.t0 is not None and self.t1 is not None else None
        post = (self.p0, self.p1) if self.p0 is not None and self.p1 is not None else None
        return summarise_clock_rows(rows, window, post)


def post_roll(sampler, step, seconds_per_step, need=0.3, length=0.6):
    """Extra untimed steps under the same load when the timed region (< `need` seconds) was too short for clock samples.
    Every rank calls it with the same (max-reduced) numbers, so all run the same number of collective steps."""
    timed = sampler.timed_seconds()
    if timed is None or timed >= need or seconds_per_step <= 0:
        return 0
    n = max(1, min(200, int(length / seconds_per_step) + 1))
    sampler.post_begin()
    for _ in range(n):
        step()
    import torch
    torch.cuda.synchronize()
    sampler.post_end()
    return n


def measured_peak():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"


def measured_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the latest committed
    `ncu --set full` capture (profiles/rNN_traffic.json); only the default workload has one."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if workload != "c4" or not files:
        return None, None
    try:
        ent = json.load(open(files[-1])).get(kernel)
        return (float(ent["traffic"]), ent["source"]) if ent else (None, None)
    except Exception:
        return None, None


def calibrate(cpu_ref, A, B, k):
    """Seconds per CG iteration of the oracle on this host (after one warm call)."""
    cpu_ref.lib().cpu_ref_set_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1
    cpu_ref.cg(A.data, A.indptr, A.indices, B, k=k, iters=1)
    t0 = time.perf_counter()
    cpu_ref.cg(A.data, A.indptr, A.indices, B, k=k, iters=7)
    return max((time.perf_counter() - t0) / 8.0, 1e-7)          # 7 iterations + the initialisation


def cpu_sample(A, B, k, target_s, max_iters):
    """Time the oracle port on the host cores for about target_s seconds of CG iterations."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_ref
    cpu_ref.build()
    per_it = calibrate(cpu_ref, A, B, k)
    iters = int(max(4, min(max_iters, target_s / per_it)))
    t0 = time.perf_counter()
    cpu_ref.cg(A.data, A.indptr, A.indices, B, k=k, iters=iters)
    dt = time.perf_counter() - t0
    return iters, dt, cpu_ref.threads()


def numpy_sample(A, B, k, target_s):
    """oracle/np_cg.cg (the restatement of helmFE_var.CG, numpy/scipy, one thread) on the first right-hand side,
    for about target_s seconds: iterations/s of ONE column, scaled to k columns solved one after the other."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import np_cg
    from threadpoolctl import threadpool_limits
    n = A.shape[0]
    b = np.ascontiguousarray(B[:n]).astype(complex if np.iscomplexobj(B) else float)
    with threadpool_limits(limits=1):                           # "single thread" includes the BLAS behind np.dot
        t0 = time.perf_counter()
        np_cg.cg(A, b, x=np.zeros(n, dtype=b.dtype), maxit=1)
        per_it = max(time.perf_counter() - t0, 1e-6) / 2.0      # initial residual + one iteration
        iters = int(max(2, min(ITERS_PER_STEP, target_s / per_it)))
        t0 = time.perf_counter()
        np_cg.cg(A, b, x=np.zeros(n, dtype=b.dtype), maxit=iters)
        dt = time.perf_counter() - t0
    return {"value": iters / dt, "unit": "iterations/s", "cores": 1,
            "sample": f"{iters} iterations of oracle/np_cg.cg (numpy/scipy) on one right-hand side, {dt:.1f} s"}


def run_reference(args, wl, dtype, k, rank, world):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the OpenCL
    original cannot run in this image: no ICD); rank 0 only."""
    if rank != 0:
        return
    A, B = make_problem(args.workload, dtype, k)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_ref
    cpu_ref.build()
    per_it = calibrate(cpu_ref, A, B, k)
    # a step is a bounded sample: as many of the 256 iterations as fit ~2 s of host time
    budget = 120.0 / max(1, args.steps + args.warmup)
    iters = int(max(2, min(ITERS_PER_STEP, min(2.0, budget) / per_it)))
    for _ in range(args.warmup):
        cpu_ref.cg(A.data, A.indptr, A.indices, B, k=k, iters=iters)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_ref.cg(A.data, A.indptr, A.indices, B, k=k, iters=iters)
    dt = time.perf_counter() - t0
    value = args.steps * iters * k / dt
    line = {
        "impl": "reference", "metric": "CG iters/sec", "value": value, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": config_of(args, wl, A, k, dtype, world),
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cpu_ref.threads(), "kind": "port",
                         "sample": f"{iters} of the {ITERS_PER_STEP} iterations of a step, x {args.steps} steps, "
                                   f"oracle/cpu_ref.c (OpenMP) on {os.cpu_count()} host cpus"},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def config_of(args, wl, A, k, dtype, world, n=None, nnz=None):
    v = A.dtype.itemsize
    per_gpu_rows = A.shape[0] // world if (world > 1 and args.mode == "row-block") else A.shape[0]
    matrix_bytes = A.nnz * (v + 4) // (world if args.mode == "row-block" else 1)
    vec_bytes = 6 * per_gpu_rows * k * v          # x, q and the two buffers each of r and d (k = 1), per GPU
    return {"workload": f"{wl['desc']}; n={A.shape[0]}, nnz={A.nnz}, k={k} per GPU, dtype {dtype}; "
                        f"step = one cg() call of {ITERS_PER_STEP} iterations from x0=0",
            "name": args.workload, "n": int(A.shape[0]), "nnz": int(A.nnz), "k": int(k),
            "iters_per_step": ITERS_PER_STEP,
            "parallelism": "single GPU" if world == 1 else
                           (f"rhs-split x{world} (matrix replicated, one RHS per GPU, no collective)" if args.mode == "rhs-split"
                            else f"row-block x{world}: the boundary entries of d and r are stored into the peers' halos over "
                                 f"NVLink by the kernels that produce them, the 2 dot products per iteration are all-reduced "
                                 f"inside the kernels through peer memory"),
            "l2": (f"no L2 flush: the vectors an iteration touches ({vec_bytes / 1e6:.0f} MB per GPU, + {matrix_bytes / 1e6:.0f} MB of "
                   f"CSR arrays when the row-pattern dictionary is off) exceed the 126 MB L2"
                   if vec_bytes > 126e6 or matrix_bytes > 126e6 else
                   f"working set per GPU ({vec_bytes / 1e6:.0f} MB of vectors) fits the 126 MB L2 (latency-bound config); "
                   f"no flush between iterations")}


def save_trace(args, M, world, rank):
    """--opt trace=N: keep the kernels' own timeline of the last step (gpurun_out/trace_<workload>_n<world>_r<rank>.npy)."""
    n = [int(o.split("=")[1]) for o in args.opt if o.startswith("trace=")]
    if n and n[0] > 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.save(os.path.join(ROOT, "gpurun_out", f"trace_{args.workload}_n{world}_r{rank}.npy"),
                M.read_trace(min(n[0], ITERS_PER_STEP)))


def kernel_specs(M, n, nnz, k, dtype):
    """[(kernel name, algorithmic bytes, bytes really moved)] of the iteration the handle runs for k right-hand sides."""
    import cg_b200.problems as P
    v = P.DTYPES[dtype][2]
    b_spmv, _ = P.algorithmic_bytes(n, nnz, k, dtype)
    npat = M.get_option("patterns") if (k == 1 and M.get_option("pattern")) else 0
    if k == 1 and npat > 0 and M.get_option("cg2_ok") and M.get_option("cg2") and M.get_option("solver") != 2:
        # two kernels per iteration (csrc/cg2.cuh): algorithmic = the SURVEY 8(d) bytes of the operations each one
        # replaces (spmv + aypx + the x-axpy | the r-axpy + vdot); moved = what the format really has to move
        return [("dir_spmv", b_spmv + 6 * n * v, n * (2 + 6 * v)), ("update_r", 3 * n * v, 3 * n * v)], npat
    moved_spmv = n * (2 + 2 * v) if npat > 0 else b_spmv
    return [("spmv_dot", b_spmv, moved_spmv), ("update_xr", 6 * k * n * v, 6 * k * n * v),
            ("update_d", 3 * k * n * v, 3 * k * n * v)], npat


def time_kernels(M, specs, k, peak, reps):
    out = {}
    for name, nbytes, moved in specs:
        kms = M.time_kernel(name, k=k, reps=reps)
        out[name] = {"ms": kms, "algorithmic_bytes": nbytes, "gbs": nbytes / kms / 1e6, "frac": nbytes / kms / 1e6 / peak,
                     "moved_bytes": moved, "moved_gbs": moved / kms / 1e6, "moved_frac": moved / kms / 1e6 / peak}
    return out


def also_single_gpu(name, dtype, peak, tdt_of, stream, k=1, opts=None, steps=5):
    """Short resident-data measurement of another BASELINE config on this GPU (value + kernel rooflines)."""
    import torch
    import cg_b200
    import cg_b200.problems as P
    A, B = make_problem(name, dtype, k)
    n, nnz = A.shape[0], A.nnz
    _, b_iter = P.algorithmic_bytes(n, nnz, k, dtype)
    M = cg_b200.Matrix.from_scipy(A)
    M.set_stream(stream.cuda_stream)
    for key, val in (opts or {}).items():
        M.set_option(key, val)
    with torch.cuda.stream(stream):
        b_dev = torch.from_numpy(B).to("cuda")
        x_dev = torch.zeros(n * k, dtype=tdt_of[dtype], device="cuda")
        for _ in range(3):
            x_dev.zero_()
            M.solve(b_dev, x=x_dev, k=k, max_iterations=ITERS_PER_STEP)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            x_dev.zero_()
            info = M.solve(b_dev, x=x_dev, k=k, max_iterations=ITERS_PER_STEP)[1]
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out = {"workload": WORKLOADS[name]["desc"], "n": int(n), "nnz": int(nnz), "k": int(k), "dtype": dtype,
               "options": opts or {}, "value": k * ITERS_PER_STEP / (ms / 1e3), "unit": "iterations/s", "ms_per_step": ms}
        it_ms = info.timing_ms["iterations"] / ITERS_PER_STEP
        out["iteration"] = {"ms": it_ms, "gbs": b_iter / it_ms / 1e6, "frac": b_iter / it_ms / 1e6 / peak}
        if M.get_option("solver") != 2 and (b_iter > 40e6 or M.get_option("solver") == 1):
            specs, npat = kernel_specs(M, n, nnz, k, dtype)
            out["kernels"] = time_kernels(M, specs, k, peak, 100 if b_iter < 500e6 else 30)
            if npat:
                out["format"] = f"row-pattern dictionary, {npat} distinct rows"
        else:
            out["kernels"] = "single cooperative launch for the whole solve (working set fits the L2): no per-kernel timing"
    M.close()
    return out


# relative recursive residual sqrt(|delta_256| / |delta_0|) of a full step (256 iterations from x0 = 0), as the
# single-GPU engine and the CPU oracle compute it (tests/test_gpu_cg2.py::test_config4_*): what every sharded run must reproduce
EXPECTED_RELRES_256 = {("c4", "f64"): 0.16825512278}


def roofline_of(kernels, peak, peak_src, traffic, traffic_src, where=""):
    """The `roofline` object for the kernel that takes the most time.  `achieved` / `frac` are on ALGORITHMIC bytes
    (SURVEY.md 8(d)); a kernel that runs from the row-pattern dictionary moves fewer bytes than that, so its fraction
    on the bytes it really moves is given too (moved_*)."""
    dom = max(kernels, key=lambda nm: kernels[nm]["ms"])
    kd = kernels[dom]
    out = {"bound": "hbm", "kernel": dom + where, "achieved": kd["gbs"], "peak": peak, "unit": "GB/s", "frac": kd["frac"],
           "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
           "algorithmic_bytes_per_launch": kd["algorithmic_bytes"], "ms_per_launch": kd["ms"]}
    if kd.get("moved_bytes", kd["algorithmic_bytes"]) != kd["algorithmic_bytes"]:
        out.update(moved_bytes_per_launch=kd["moved_bytes"], moved_achieved=kd["moved_gbs"], moved_frac=kd["moved_frac"],
                   note="this kernel runs from the row-pattern dictionary (2 bytes per row instead of the CSR arrays): "
                        "`frac` on the CSR-based algorithmic bytes can exceed 1 -- that is format compression; "
                        "`moved_frac` is the fraction of the HBM roofline on the bytes it really moves")
    return out


def sharded_parity(args, M, plan, b_owned, rb, re, dtype, rank, world, dist, torch, prefix_iters=8):
    """Collective.  (1) `prefix_iters` iterations of the sharded solve, x gathered on rank 0 and compared with
    oracle/cpu_ref.c run on the WHOLE system there; (2) the recursive residual after a full 256-iteration step
    against the single-GPU value.  Returns the `parity` object of the JSON line (rank 0; a dict with ok on all)."""
    import cg_b200.problems as P
    np_t = P.DTYPES[dtype][0]
    x, info = M.solve(b_owned, np.zeros(re - rb, dtype=np_t), max_iterations=prefix_iters)
    sizes = [int(plan.bounds[p + 1] - plan.bounds[p]) for p in range(world)]
    xt = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    outs = [xt if p == rank else torch.zeros(sizes[p], dtype=xt.dtype, device="cuda") for p in range(world)]
    for p in range(world):                      # (slices differ in length: one broadcast per owner)
        dist.broadcast(outs[p], src=p)
    _, full = M.solve(b_owned, np.zeros(re - rb, dtype=np_t), max_iterations=ITERS_PER_STEP)
    res = {"ok": True}
    if rank == 0:
        xg = torch.cat(outs).cpu().numpy()
        A, B = make_problem(args.workload, dtype, 1)
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import cpu_ref
        cpu_ref.build()
        cpu_ref.lib().cpu_ref_set_threads(os.cpu_count() or 1)
        ref, _, _ = cpu_ref.cg(A.data, A.indptr, A.indices, B, k=1, iters=prefix_iters)
        err = float(np.linalg.norm(xg - ref) / np.linalg.norm(ref))
        bar = 1e-10 if dtype in ("f64", "c128") else 1e-5
        res = {"max_rel": err, "bar": bar, "iters": prefix_iters, "against": "oracle/cpu_ref.c on the whole system (rank 0 host)",
               "relres_256": full["relres"], "ok": bool(err < bar)}
        exp = EXPECTED_RELRES_256.get((args.workload, dtype))
        if exp is not None:
            res["relres_256_single_gpu"] = exp
            res["ok"] = bool(res["ok"] and abs(full["relres"] - exp) < 1e-8 * exp)
    flag = torch.tensor([1 if res["ok"] else 0], device="cuda")
    dist.broadcast(flag, src=0)
    if rank != 0:
        res = {"ok": bool(flag.item())}
    return res


def run_row_block(args, wl, dtype, rank, local_rank, world):
    """N > 1: the CSR matrix row-block partitioned over the ranks (strong scaling of ONE system)."""
    import torch
    import torch.distributed as dist
    import cg_b200
    import cg_b200.problems as P
    from cg_b200 import sharded
    np_t, _, v_bytes, cplx = P.DTYPES[dtype]
    if args.workload in ("c3", "c4", "c4slab4", "c4slab8"):
        N3 = 128 if args.workload == "c3" else 300
        NZ = {"c4slab4": 76, "c4slab8": 38}.get(args.workload, N3)
        n = N3 * N3 * NZ
        planes = [(NZ * p) // world for p in range(world + 1)]          # whole z-planes per rank
        bounds = np.array([pl * N3 * N3 for pl in planes], dtype=np.int64)
        rb, re = int(bounds[rank]), int(bounds[rank + 1])
        Al = P.laplace3d(N3, dtype=np_t, rows=(rb, re), nz=NZ)
        nnz = 7 * n - 2 * N3 * N3 - 4 * N3 * NZ
        plan = sharded.plan_row_block(Al.indptr, Al.indices, Al.data, bounds, rank)
        b_owned = np.ones(re - rb, dtype=np_t)
        shape = (n, nnz)
    else:
        A, B = make_problem(args.workload, dtype, 1)
        n, nnz = A.shape[0], A.nnz
        bounds = sharded.split_rows(A.indptr, world, by="nnz" if args.workload == "c5" else "rows")
        rb, re = int(bounds[rank]), int(bounds[rank + 1])
        plan = sharded.plan_row_block(A.indptr, A.indices, A.data, bounds, rank)
        b_owned = np.ascontiguousarray(B[rb:re])
        shape = (n, nnz)
        del A
    sampler = ClockSampler(local_rank)          # (early: nvidia-smi needs ~0.5 s to come up; samples are filtered by time)
    M = sharded.ShardedMatrix(plan, device=local_rank)
    stream = torch.cuda.Stream()
    M.set_stream(stream.cuda_stream)
    for kv in args.opt:
        key, val = kv.split("=")
        M.set_option(key, int(val))
    tdt = {"f32": torch.float32, "f64": torch.float64, "c64": torch.complex64, "c128": torch.complex128}[dtype]

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        b_dev = torch.from_numpy(b_owned).to("cuda")
        x_dev = torch.zeros(re - rb, dtype=tdt, device="cuda")

        def step():
            x_dev.zero_()
            return M.solve(b_dev, x_dev, max_iterations=ITERS_PER_STEP)[1]

        for _ in range(args.warmup):
            step()
        l0 = M.info()["launches"]
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.begin()
        e0.record(stream)
        for _ in range(args.steps):
            info = step()
        e1.record(stream)
        barrier()
        sampler.end()
        ms = e0.elapsed_time(e1)
        launches = M.info()["launches"] - l0
        save_trace(args, M, world, rank)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item()) / args.steps
        # (the same decision on every rank: the wall-clock length of the region is reduced too)
        tw = torch.tensor([sampler.timed_seconds() or 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        sampler.t1 = sampler.t0 + float(tw.item())
        post_roll(sampler, step, ms_per_step / 1e3)
        clocks = sampler.stop()
    value = ITERS_PER_STEP / (ms_per_step / 1e3)

    # e2e: this rank's slices of b and x0 in pinned host memory, x read back, every step
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    b_h, x_h = pin(b_owned), pin(np.zeros_like(b_owned))
    def e2e_step():
        x_h[...] = 0
        dist.barrier()
        t0 = time.perf_counter()
        M.solve(b_h, x_h, max_iterations=ITERS_PER_STEP)
        return time.perf_counter() - t0
    e2e_step()
    tt = sum(e2e_step() for _ in range(args.steps))
    t = torch.tensor([tt], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tt = float(t.item())
    tot = torch.tensor([float(plan.n_owned * v_bytes)], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)
    e2e = {"value": ITERS_PER_STEP * args.steps / tt, "unit": "iterations/s",
           "h2d_bytes_per_step": int(2 * tot.item()), "d2h_bytes_per_step": int(tot.item()),
           "ms_per_step": 1e3 * tt / args.steps,
           "how": "cgb200_shard_solve on pinned host slices of b / x0 per rank, x read back; the row blocks of the "
                  "matrix stay resident (they are placed once by cgb200_shard_create)"}

    peak, peak_src = measured_peak()
    ln, lnnz = plan.n_owned, int(plan.data.size)
    b_spmv = lnnz * (v_bytes + 4) + 4 * (ln + 1) + 2 * ln * v_bytes
    kernels = {}
    reps = 50 if b_spmv > 50e6 else 400
    npat = M.get_option("patterns")
    two_kernel = bool(M.get_option("cg2_ok") and M.get_option("cg2") and npat > 0)
    if two_kernel:
        # two kernels per iteration (csrc/cg2.cuh).  algorithmic = the SURVEY 8(d) bytes of the operations each one
        # replaces (spmv + aypx + the x-axpy | the r-axpy + vdot); moved = what the format really has to move
        specs = (("dir_spmv", b_spmv + 6 * ln * v_bytes, ln * (2 + 6 * v_bytes)), ("update_r", 3 * ln * v_bytes, 3 * ln * v_bytes))
    else:
        moved_spmv = ln * (2 + 2 * v_bytes) if npat > 0 else b_spmv
        specs = (("spmv_dot", b_spmv, moved_spmv), ("update_xr", 6 * ln * v_bytes, 6 * ln * v_bytes),
                 ("update_d", 3 * ln * v_bytes, 3 * ln * v_bytes))
    for name, nbytes, moved in specs:
        kms = M.time_kernel(name, reps=reps)
        kernels[name] = {"ms": kms, "algorithmic_bytes": nbytes, "gbs": nbytes / kms / 1e6, "frac": nbytes / kms / 1e6 / peak,
                         "moved_bytes": moved, "moved_gbs": moved / kms / 1e6, "moved_frac": moved / kms / 1e6 / peak}
    if npat > 0:
        kernels[specs[0][0]]["format"] = f"row-pattern dictionary, {npat} distinct rows (DESIGN.md 4.3)"
    sinfo = M.info()

    # ---- parity, outside the timed region: a prefix of the iteration against the CPU oracle on the WHOLE system
    # (rank 0 builds it), and the recursive residual after a full step against the single-GPU engine's value
    parity = sharded_parity(args, M, plan, b_owned, rb, re, dtype, rank, world, dist, torch)
    M.close()
    if parity is not None and not parity["ok"]:
        if rank == 0:
            print("bench.py: PARITY FAILURE of the sharded solve: " + json.dumps(parity), file=sys.stderr, flush=True)
        dist.destroy_process_group()
        raise SystemExit(3)
    if rank == 0:
        it_ms = info["timing_ms"]["iterations"] / ITERS_PER_STEP

        class _A:  # shape carrier for config_of
            pass
        a = _A()
        a.shape, a.nnz, a.dtype = (shape[0], shape[0]), shape[1], np.dtype(np_t)
        b_iter_local = b_spmv + 9 * ln * v_bytes
        line = {
            "metric": "CG iters/sec", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": config_of(args, wl, a, 1, dtype, world),
            "roofline": roofline_of(kernels, peak, peak_src, None, None, where=" (rank 0 row block)"),
            "kernels": kernels, "parity": parity,
            "iteration": {"ms": it_ms, "local_algorithmic_bytes": b_iter_local, "gbs_per_gpu": b_iter_local / it_ms / 1e6,
                          "frac": b_iter_local / it_ms / 1e6 / peak},
            "shard": sinfo, "options": args.opt, "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "solve_phases_ms": info["timing_ms"],
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-also", action="store_true", help="N = 1: skip the extra C2 (Helmholtz, complex) figures")
    ap.add_argument("--dtype", default=None, choices=["f32", "f64", "c64", "c128"])
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--mode", default="row-block", choices=["row-block", "rhs-split"],
                    help="N > 1: row-block sharding with halo exchange + all-reduce (default), or the reference's "
                         "rhs-split (matrix replicated, no collective)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (cgb200_set_option)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    wl = WORKLOADS[args.workload]
    dtype = args.dtype or wl["dtype"]
    k = args.k or wl["k"]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, wl, dtype, k, rank, world)
        return

    import torch
    import torch.distributed as dist
    import cg_b200
    import cg_b200.problems as P
    if not torch.cuda.is_available() or cg_b200.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if world > 1 and args.mode == "row-block" and k == 1:
        run_row_block(args, wl, dtype, rank, local_rank, world)
        return

    A, B = make_problem(args.workload, dtype, k)
    if world > 1 and k == 1:
        # rhs-split: every rank gets its own right-hand side (rank 0 keeps the config's b)
        if rank > 0:
            v = np.random.default_rng(1000 + rank).uniform(-1.0, 1.0, A.shape[0])
            B = np.ascontiguousarray(v + 0j if P.DTYPES[dtype][3] else v, dtype=A.dtype)
    n, nnz = A.shape[0], A.nnz
    v_bytes = P.DTYPES[dtype][2]
    b_spmv, b_iter = P.algorithmic_bytes(n, nnz, k, dtype)

    stream = torch.cuda.Stream()
    tdt_of = {"f32": torch.float32, "f64": torch.float64, "c64": torch.complex64, "c128": torch.complex128}
    tdt = tdt_of[dtype]
    M = cg_b200.Matrix.from_scipy(A, device=local_rank)
    M.set_stream(stream.cuda_stream)
    for kv in args.opt:
        key, val = kv.split("=")
        M.set_option(key, int(val))
    with torch.cuda.stream(stream):
        b_dev = torch.from_numpy(B).to("cuda", non_blocking=False)
        x_dev = torch.zeros(n * k, dtype=tdt, device="cuda")

    def step():
        x_dev.zero_()
        M.solve(b_dev, x=x_dev, k=k, max_iterations=ITERS_PER_STEP, tol=0.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)          # nvidia-smi needs ~0.5 s to start: the warm-up steps cover it
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
        launches0 = M.info()["launches"]
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.begin()
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
        sampler.end()
        ms = e0.elapsed_time(e1)
        launches = M.info()["launches"] - launches0
        save_trace(args, M, world, rank)
        if world == 1:
            post_roll(sampler, step, ms / 1e3 / max(1, args.steps))     # (tiny configs: a step is a few milliseconds)
        clocks = sampler.stop()
        timing = M.solve(b_dev, x=x_dev, k=k, max_iterations=ITERS_PER_STEP)[1].timing_ms
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * k * ITERS_PER_STEP / (ms_per_step / 1e3)    # right-hand-side iterations per second, whole job

    # ---- per-kernel rooflines (rank 0), CUDA events around back-to-back launches of one kernel
    peak, peak_src = measured_peak()
    kernels = {}
    npat = 0
    if rank == 0:
        specs, npat = kernel_specs(M, n, nnz, k, dtype)
        kernels = time_kernels(M, specs, k, peak, 50 if b_iter > 50e6 else 400)
        M.solve(b_dev, x=x_dev, k=k, max_iterations=2)          # restore a sane state
        if npat > 0:
            kernels[specs[0][0]]["format"] = f"row-pattern dictionary, {npat} distinct rows (DESIGN.md 4.3)"
    it_ms = timing["iterations"] / ITERS_PER_STEP

    # ---- e2e through the exported cg / cgd symbol with pinned HOST buffers.  Top level: the matrix stays resident between
    # calls (the as_prec pattern the symbol exists for, and the same kind of number as the sharded line's e2e at
    # N > 1): every timed call hashes the CSR arrays it is handed (content identity), uploads b and x0 and reads x
    # back.  Beside it: the matrix uploaded inside every call too, and the reference's allocate-everything life cycle.
    e2e = None
    if not args.no_e2e:
        os.environ["CGB200_DEVICE"] = str(local_rank)
        pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
        vals_h, ptr_h, cols_h = pin(A.data), pin(A.indptr.astype(np.intc)), pin(A.indices.astype(np.intc))
        b_h, x_h = pin(B), pin(np.zeros_like(B))

        def e2e_step():
            x_h[...] = 0
            t0 = time.perf_counter()
            cg_b200.cg(n, nnz, vals_h, b_h, ptr_h, cols_h, x_h, k, ITERS_PER_STEP)
            return time.perf_counter() - t0

        def e2e_run(mode):
            os.environ["CGB200_CACHE"] = mode
            for _ in range(2):
                e2e_step()
            barrier()
            tt = sum(e2e_step() for _ in range(args.steps))
            t = torch.tensor([tt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        its_total = world * k * ITERS_PER_STEP * args.steps
        h2d_matrix = nnz * (v_bytes + 4) + 4 * (n + 1)
        tt1 = e2e_run("1")
        e2e = {"value": its_total / tt1, "unit": "iterations/s",
               "h2d_bytes_per_step": int(2 * k * n * v_bytes), "d2h_bytes_per_step": int(k * n * v_bytes),
               "ms_per_step": 1e3 * tt1 / args.steps, "kind": "matrix resident",
               "how": "cg()/cgd() C symbol on pinned host numpy arrays, wall clock around the blocking call; the matrix "
                      "of the previous call is still resident (default CGB200_CACHE=1: the call hashes the CSR arrays "
                      "it is handed -- 128-bit content identity -- and re-uploads only on a change); b and x0 are "
                      "uploaded and x is read back inside every timed call"}
        tt2 = e2e_run("2")          # device buffers reused, content uploaded on every call
        e2e["matrix_uploaded"] = {"value": its_total / tt2, "ms_per_step": 1e3 * tt2 / args.steps,
                                  "h2d_bytes_per_step": int(h2d_matrix + 2 * k * n * v_bytes)}
        # the reference's own life cycle: allocate, upload, solve, free on every call (clcg.c:142-214, :432-459)
        tt0 = e2e_run("0")
        e2e["alloc_every_call"] = {"value": its_total / tt0, "ms_per_step": 1e3 * tt0 / args.steps}
        cg_b200._lib.lib().cgb200_clear_cache()

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            iters, dt, threads = cpu_sample(A, B, k, 12.0, ITERS_PER_STEP)
            cpu = {"value": iters * k / dt, "unit": "iterations/s", "cores": threads, "kind": "port",
                   "sample": f"{iters} CG iterations of the same system on the host "
                             f"(oracle/cpu_ref.c, OpenMP, {threads} threads of {os.cpu_count()} cpus), {dt:.1f} s"}
            try:        # second point of BASELINE.md section 3: the reference's numpy CG recurrence, one thread
                cpu["numpy_single_thread"] = numpy_sample(A, B, k, 6.0)
            except Exception as e:      # never let a reporting extra break the line
                cpu["numpy_single_thread"] = {"error": str(e)[:100]}
        dom = max(kernels, key=lambda nm: kernels[nm]["ms"])
        traffic, traffic_src = measured_traffic(args.workload if (dtype, k) == (wl["dtype"], wl["k"]) else None, dom)
        line = {
            "metric": "CG iters/sec", "value": value, "unit": "iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": config_of(args, wl, A, k, dtype, world),
            "roofline": roofline_of(kernels, peak, peak_src, traffic, traffic_src),
            "kernels": kernels,
            "iteration": {"ms": it_ms, "algorithmic_bytes": b_iter, "gbs": b_iter / it_ms / 1e6,
                          "frac": b_iter / it_ms / 1e6 / peak,
                          "kernels_per_iteration": len(kernels)},
            "options": args.opt, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "solve_phases_ms": timing,
        }
        if npat > 0:
            line["iteration"]["note"] = ("the SpMV runs from the row-pattern dictionary and moves fewer bytes than the CSR "
                                         "figure the algorithmic bytes are (kernels.*.moved_*): fractions above 1 are format "
                                         "compression, not bandwidth")
        if world == 1 and not args.no_also:
            # the other BASELINE configs on this GPU, resident data, short; c4_csr = the CSR kernels (the north star's
            # SpMV: nnz(v+4) bytes streamed by TMA) on the headline system, with the row-pattern dictionary switched off
            also = {}
            legs = [("c4_csr", "c4", "f64", 1, {"pattern": 0}, 3), ("c2_c128", "c2", "c128", 1, None, 5),
                    ("c1_f64", "c1", "f64", 1, None, 5), ("c3_f64_k32", "c3", "f64", 32, None, 3),
                    ("c5_f64", "c5", "f64", 1, None, 3)]
            for key, name, dt_, k_, opts, steps_ in legs:
                if name == args.workload and not opts:
                    continue
                try:
                    also[key] = also_single_gpu(name, dt_, peak, tdt_of, stream, k=k_, opts=opts, steps=steps_)
                except Exception as e:          # a reporting extra must never break the line
                    also[key] = {"error": str(e)[:200]}
            line["also"] = also
        print(json.dumps(line), flush=True)
    M.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
