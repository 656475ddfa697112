#!/bin/bash
# gpurun call 2 of session 2: the balanced TMA-stream kernel on C5, and what makes PDL slow inside the iteration.
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
B="timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-also --steps 3"

timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest13.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest13.log
tail -3 $O/pytest13.log

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

run c5_auto --workload c5 --opt pdl=0
run c5_v4 --workload c5 --opt pdl=0 --opt spmv_variant=4
run c5_v5 --workload c5 --opt pdl=0 --opt spmv_variant=5
run c5_rows32 --workload c5 --opt pdl=0 --opt auto_irregular=0 --opt defer_len=32
for w in c2 c4slab8; do
run ${w}_p0 --workload $w --opt pdl=0
run ${w}_p7 --workload $w --opt pdl=7
run ${w}_p7late --workload $w --opt pdl=7 --opt pdl_early=0
run ${w}_p1 --workload $w --opt pdl=1
run ${w}_p2 --workload $w --opt pdl=2
run ${w}_p4 --workload $w --opt pdl=4
run ${w}_p7cv --workload $w --opt pdl=7 --opt vec_carveout=100
run ${w}_p0cv --workload $w --opt pdl=0 --opt vec_carveout=100
run ${w}_p7nograph --workload $w --opt pdl=7 --opt use_graph=0
run ${w}_p0nograph --workload $w --opt pdl=0 --opt use_graph=0
done
run c4_p1 --workload c4 --opt pdl=1
run c4_p7late --workload c4 --opt pdl=7 --opt pdl_early=0
