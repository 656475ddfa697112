#!/usr/bin/env python3
"""Turn an ncu report into the two text files kept under profiles/ and one entry of rNN_traffic.json.

    python tools/ncu_summary.py gpurun_out/r02_dir_march_c4.ncu-rep --name dir_spmv [--algorithmic BYTES] [--rm]

Writes <stem>.txt (the `--page details` sections that matter + the hottest source lines by stall samples) and
<stem>.csv (`--page raw --csv`, every metric of the launch), prints the traffic entry as JSON.  `.ncu-rep` files are
tens of MB each and gpurun returns at most 64 MiB per call: --rm deletes the report once the text exists."""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

KEEP = ("Duration", "Throughput", "Issue Slots", "Executed Ipc", "Warp Cycles Per Issued", "No Eligible", "Eligible Warps",
        "Active Warps", "Hit Rate", "Registers Per Thread", "Shared Memory", "Block Size", "Grid Size", "Mem Busy", "Mem Pipes",
        "Max Bandwidth", "Theoretical", "Achieved", "Section:", "Context", "DRAM", "L1/TEX", "L2 ")


def run(*a):
    return subprocess.run(["ncu", *a], capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--name", required=True)
    ap.add_argument("--algorithmic", type=float, default=None)
    ap.add_argument("--rm", action="store_true")
    args = ap.parse_args()
    stem = args.rep[:-len(".ncu-rep")]
    raw = run("-i", args.rep, "--page", "raw", "--csv")
    open(stem + ".csv", "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    d = dict(zip(rows[0], rows[2])) if len(rows) > 2 else {}
    unit = dict(zip(rows[0], rows[1])) if len(rows) > 1 else {}

    def val(k, to="byte"):
        try:
            v = float(d[k].replace(",", ""))
        except Exception:
            return None
        u = unit.get(k, "")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
        return v * scale
    det = run("-i", args.rep, "--page", "details")
    lines = [l for l in det.splitlines() if any(k in l for k in KEEP)]
    stalls = sorted(((float(v), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
                     and v not in ("", "n/a")), reverse=True)
    src = run("-i", args.rep, "--page", "source", "--csv", "--print-source", "cuda,sass")
    hot, cur, hdr = [], None, None
    for r in csv.reader(io.StringIO(src)):
        if r and r[0] == "File Path":
            cur = r[1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and cur and len(r) >= 8:
            e = dict(zip(hdr, r))
            try:
                hot.append((int(e["# Samples"] or 0), int(e["Instructions Executed"] or 0),
                            int(float(e.get("L1 Wavefronts Shared") or 0)), os.path.basename(cur), int(e["Line No"])))
            except Exception:
                pass
    ts, ti = sum(h[0] for h in hot) or 1, sum(h[1] for h in hot) or 1
    with open(stem + ".txt", "w") as f:
        f.write(f"# {os.path.basename(args.rep)}: ncu --set full --clock-control none --import-source on, one launch (cold caches, serialised)\n")
        f.write("\n".join(lines) + "\n\n# warp stall reasons, cycles per issued instruction\n")
        f.write("\n".join(f"{k:28s} {v:6.2f}" for v, k in stalls[:12]) + "\n")
        for k in ("smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
                  "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
                  "lts__t_bytes.sum", "gpu__time_duration.sum"):
            if k in d:
                f.write(f"{k} = {d[k]} {unit.get(k, '')}\n")
        f.write("\n# hottest source lines: % of stall samples, % of executed instructions, shared-memory wavefronts (M)\n")
        for h in sorted(hot, reverse=True)[:25]:
            f.write(f"{h[3]}:{h[4]:<5d} {100 * h[0] / ts:5.1f} %  {100 * h[1] / ti:5.1f} %  {h[2] / 1e6:7.2f}\n")
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    entry = {args.name: {"kernel": d.get("Kernel Name"), "dram_bytes_read": rd, "dram_bytes_write": wr,
                         "traffic": (rd + wr) if rd is not None and wr is not None else None,
                         "algorithmic_bytes": args.algorithmic, "duration_us_under_ncu": val("gpu__time_duration.sum"),
                         "source": f"profiles/{os.path.basename(stem)}.csv (ncu --set full --clock-control none, one launch)"}}
    print(json.dumps(entry))
    if args.rm:
        os.remove(args.rep)


if __name__ == "__main__":
    main()
