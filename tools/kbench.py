#!/usr/bin/env python3
"""Kernel-level A/B timing on one GPU: the problem is generated and uploaded ONCE, then every option set of
--set runs 256-iteration solves (CUDA events from the engine) and the loop kernels alone.

    python tools/kbench.py --workload c4 --set cg2=1 --set cg2=0 --set cg2=1,cg2_blocks=2
One JSON line per option set on stdout."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (problem generators, byte counts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--dtype", default=None)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--iters", type=int, default=256)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--set", action="append", default=[], help="comma-separated key=value engine options")
    args = ap.parse_args()
    import torch
    import cg_b200
    import cg_b200.problems as P
    wl = bench.WORKLOADS[args.workload]
    dtype, k = args.dtype or wl["dtype"], args.k or wl["k"]
    A, B = bench.make_problem(args.workload, dtype, k)
    n, nnz = A.shape[0], A.nnz
    v = P.DTYPES[dtype][2]
    peak, _ = bench.measured_peak()
    b_spmv, b_iter = P.algorithmic_bytes(n, nnz, k, dtype)
    tdt = {"f32": torch.float32, "f64": torch.float64, "c64": torch.complex64, "c128": torch.complex128}[dtype]
    M = cg_b200.Matrix.from_scipy(A)
    b_dev = torch.from_numpy(B).cuda()
    x_dev = torch.zeros(n * k, dtype=tdt, device="cuda")
    for spec in (args.set or [""]):
        opts = dict(kv.split("=") for kv in spec.split(",") if kv)
        for key, val in opts.items():
            M.set_option(key, int(val))
        two = k == 1 and M.get_option("cg2_ok") == 1 and M.get_option("cg2") == 1
        best = None
        for _ in range(args.reps + 1):
            x_dev.zero_()
            info = M.solve(b_dev, x=x_dev, k=k, max_iterations=args.iters)[1]
            torch.cuda.synchronize()
            ms = info.timing_ms["iterations"]
            best = ms if best is None else min(best, ms)
        out = {"workload": args.workload, "dtype": dtype, "k": k, "options": opts, "two_kernel": bool(two),
               "patterns": int(M.get_option("patterns")) if k == 1 else 0,
               "us_per_iteration": 1e3 * best / args.iters, "its_per_s": args.iters / best * 1e3,
               "relres": float(info.relres[0]), "kernels": {}}
        if two:
            names = {"dir_spmv": n * (2 + 6 * v), "update_r": 3 * n * v}
        else:
            moved_spmv = n * (2 + 2 * v) if out["patterns"] and opts.get("pattern", "1") != "0" else b_spmv
            names = {"spmv_dot": moved_spmv, "update_xr": 6 * k * n * v, "update_d": 3 * k * n * v}
        for nm, moved in names.items():
            kms = M.time_kernel(nm, k=k, reps=50 if b_iter > 50e6 else 300)
            out["kernels"][nm] = {"us": round(1e3 * kms, 2), "moved_bytes": moved, "moved_frac": round(moved / kms / 1e6 / peak, 3)}
        out["iteration_frac_of_algorithmic"] = round(b_iter / (best / args.iters) / 1e6 / peak, 3)
        print(json.dumps(out), flush=True)
        for key in opts:      # back to defaults for the next set
            M.set_option(key, {"cg2": 1, "pattern": 1, "pdl": 1, "use_graph": 1, "pdl_early": 1, "graph_chunk": 16,
                               "defer_len": 16, "auto_irregular": 1, "vec_carveout": -1, "march": 1}.get(key, 0))
    M.close()


if __name__ == "__main__":
    main()
