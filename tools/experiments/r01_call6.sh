#!/bin/bash
# gpurun call (1 GPU): full GPU test suite, benches with the row-pattern dictionary, the default bench line and its ncu passes.
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
B="timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-also --steps 3"
timeout 900 python -m pytest tests -m gpu -q > $O/pytest17.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest17.log
tail -4 $O/pytest17.log
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
run c2_pat --workload c2
run c2c64_pat --workload c2 --dtype c64
run slab_pat --workload c4slab8
run c1_pat --workload c1
run c4_csr --workload c4 --opt pattern=0
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
l = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("default", round(l["value"], 1), "it/s e2e", round(l["e2e"]["value"], 1), {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in l["kernels"].items()},
      "iter_us", round(l["iteration"]["ms"] * 1e3, 1), "roofline", l["roofline"]["kernel"], round(l["roofline"]["frac"], 3), "cpu", l["cpu_baseline"]["value"], l["clocks"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file $O/launches_default.csv \
    python bench.py --no-cpu-baseline --no-e2e --no-also --steps 1 > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_pattern -s 40 -c 1 -f -o $O/spmv_pattern_c4 \
    python bench.py --no-cpu-baseline --no-e2e --no-also --steps 1 > $O/ncu_pat_c4.log 2>&1; echo "ncu pattern c4 rc=$?"
