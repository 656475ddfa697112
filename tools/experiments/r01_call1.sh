#!/bin/bash
# One gpurun call of round 1 (session 2): tests, the effect of the new options, the ncu passes.
# Usage (from the repo root on the GPU box): bash tools/experiments/r01_call1.sh
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
B="timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-also --steps 3"

timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest12.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest12.log
tail -3 $O/pytest12.log

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

run c5_defer0 --workload c5 --opt defer_len=0
run c5_defer16 --workload c5
run c5_defer32 --workload c5 --opt defer_len=32
run c5_defer8 --workload c5 --opt defer_len=8
run c5_v3 --workload c5 --opt spmv_variant=3
run c4_pdl0 --workload c4 --opt pdl=0
run c4_pdl7 --workload c4
run c2_pdl0 --workload c2 --opt pdl=0
run c2_pdl7 --workload c2
run c2c64_pdl0 --workload c2 --dtype c64 --opt pdl=0
run c2c64_pdl7 --workload c2 --dtype c64
run slab_pdl0 --workload c4slab8 --opt pdl=0
run slab_pdl7 --workload c4slab8 --opt trace=256
run c3_pdl0 --workload c3 --opt pdl=0
run c3_pdl7 --workload c3
python tools/trace_report.py $O/trace_c4slab8_n1_r0.npy

# the bench line proper (all legs), then the ncu passes of the same command
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file $O/launches_default.csv \
    python bench.py --no-cpu-baseline --no-e2e --no-also --steps 1 > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_tma_rows -s 40 -c 1 -f -o $O/spmv_c4 \
    python bench.py --no-cpu-baseline --no-e2e --no-also --steps 1 > $O/ncu_spmv_c4.log 2>&1; echo "ncu spmv c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:update_xr -s 40 -c 1 -f -o $O/xr_c4 \
    python bench.py --no-cpu-baseline --no-e2e --no-also --steps 1 > $O/ncu_xr_c4.log 2>&1; echo "ncu xr c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_tma_rows -s 40 -c 1 -f -o $O/spmv_c5 \
    python bench.py --no-cpu-baseline --no-e2e --no-also --steps 1 --workload c5 > $O/ncu_spmv_c5.log 2>&1; echo "ncu spmv c5 rc=$?"
ls -la $O/*.ncu-rep
