#!/usr/bin/env python3
"""Summarise the kernels' own timeline of a solve (bench.py --opt trace=N writes gpurun_out/trace_*.npy).

    python tools/trace_report.py gpurun_out/trace_c4_n8_r*.npy

Per iteration (median over the traced iterations, first 4 skipped), in microseconds:
  spmv        spmv_start -> all blocks done          (local work)
  spmv_red    all blocks done -> d.q known           (grid reduction + all-reduce over the GPUs: includes waiting
                                                      for the slowest GPU)
  gap1        d.q known -> x/r update starts         (kernel boundary)
  xr, xr_red, gap2 likewise; d = direction update start -> next spmv start (kernel + boundary)
"""
import sys

import numpy as np

EV = ("spmv_start", "spmv_all_done", "spmv_end", "xr_start", "xr_all_done", "xr_end", "d_start", "halo_ready")


def report(path):
    tr = np.load(path).astype(np.int64)
    two = (tr[:, :6] > 0).all(axis=1).sum() >= 8 and (tr[:, 6] == 0).all()      # two-kernel iteration: no direction kernel
    ok = (tr[:, :6 if two else 7] > 0).all(axis=1)
    tr = tr[ok]
    if len(tr) < 8:
        return f"{path}: too few complete iterations ({len(tr)})"
    a, nxt = tr[4:-1], tr[5:]
    if two:
        # dir_spmv: start -> all blocks done | -> d.q known (grid sum + all-reduce) | gap | update_r likewise | gap to the next
        seg = {
            "dir_spmv": a[:, 1] - a[:, 0], "dir_red": a[:, 2] - a[:, 1], "gap1": a[:, 3] - a[:, 2],
            "update_r": a[:, 4] - a[:, 3], "r_red": a[:, 5] - a[:, 4], "gap2": nxt[:, 0] - a[:, 5],
            "iteration": nxt[:, 0] - a[:, 0],
        }
    else:
        seg = {
            "spmv": a[:, 1] - a[:, 0], "spmv_red": a[:, 2] - a[:, 1], "gap1": a[:, 3] - a[:, 2],
            "xr": a[:, 4] - a[:, 3], "xr_red": a[:, 5] - a[:, 4], "gap2": a[:, 6] - a[:, 5],
            "d+gap3": nxt[:, 0] - a[:, 6], "iteration": nxt[:, 0] - a[:, 0],
        }
        if (a[:, 7] > 0).all():
            seg["halo_ready_after_spmv_start"] = a[:, 7] - a[:, 0]
    return path + "\n  " + "  ".join(f"{k}={np.median(v) / 1e3:.1f}" for k, v in seg.items())


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(report(p))
