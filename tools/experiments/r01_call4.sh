#!/bin/bash
# gpurun call 4 (1 GPU): L2 eviction policy for d, q, r; in-sequence DRAM traffic; ncu of the SpMM and the balanced SpMV.
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
B="timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-also --steps 3"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest15.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest15.log
tail -3 $O/pytest15.log
run() { # name, args...
    local name=$1; shift
    $B "$@" > $O/$name.json 2> $O/$name.err || echo "$name FAILED rc=$?"
    python - "$O/$name.json" "$name" <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(l["value"], 1), "it/s", {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in l["kernels"].items()},
          "iter_us", round(l["iteration"]["ms"] * 1e3, 1), round(l["iteration"]["frac"], 3))
except Exception as e:
    print(sys.argv[2], "no line", e)
PY
}
for w in c4slab8 c2; do
run ${w}_keep0 --workload $w --opt l2_keep=0
run ${w}_keep1 --workload $w --opt l2_keep=1
done
run c2c64_keep0 --workload c2 --dtype c64 --opt l2_keep=0
run c2c64_keep1 --workload c2 --dtype c64 --opt l2_keep=1
run c4_keep0 --workload c4 --opt l2_keep=0
run c4_keep1 --workload c4 --opt l2_keep=1
run c1_default --workload c1
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
for kp in 0 1; do
timeout 600 ncu --cache-control none --clock-control none --metrics $M -s 1500 -c 9 --csv --log-file $O/seq_slab_keep$kp.csv \
    python bench.py --no-cpu-baseline --no-e2e --no-also --steps 1 --workload c4slab8 --opt l2_keep=$kp > $O/seq_slab_keep$kp.log 2>&1; echo "seq slab keep$kp rc=$?"
timeout 600 ncu --cache-control none --clock-control none --metrics $M -s 1500 -c 9 --csv --log-file $O/seq_c2_keep$kp.csv \
    python bench.py --no-cpu-baseline --no-e2e --no-also --steps 1 --workload c2 --opt l2_keep=$kp > $O/seq_c2_keep$kp.log 2>&1; echo "seq c2 keep$kp rc=$?"
done
python - <<'PY'
import csv, glob
for f in sorted(glob.glob("gpurun_out/seq_*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 12 and r[0].isdigit()]
    by = {}
    for r in rows:
        by.setdefault(r[0], {"k": r[4].split("<")[0].split("::")[-1]})[r[12]] = (r[14], r[13])
    print(f)
    for i, d in by.items():
        print("  ", d["k"], {m: v for m, v in d.items() if m != "k"})
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -s 40 -c 1 -f -o $O/spmm_c3 \
    python bench.py --no-cpu-baseline --no-e2e --no-also --steps 1 --workload c3 > $O/ncu_spmm_c3.log 2>&1; echo "ncu spmm c3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_tma_kernel -s 40 -c 1 -f -o $O/spmv_c5_balanced \
    python bench.py --no-cpu-baseline --no-e2e --no-also --steps 1 --workload c5 > $O/ncu_spmv_c5b.log 2>&1; echo "ncu spmv c5 rc=$?"
run c5_default --workload c5
run c3_default --workload c3
