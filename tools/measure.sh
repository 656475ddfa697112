#!/bin/bash
# final measuring call of the session (1 GPU): tests, the bench lines kept under profiles/, the ncu passes of the default command
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
timeout 600 python -m pytest tests -m gpu -q > $O/pytest19.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest19.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
for w in c1 c2 c3 c5; do timeout 600 python bench.py --workload $w --no-also --steps 3 > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"; done
timeout 600 python bench.py --workload c2 --dtype c64 --no-also --steps 3 > $O/bench_c2_c64.json 2> $O/bench_c2_c64.err
python - <<'PY'
import json, glob
for f in ["bench_default", "bench_reference", "bench_c1", "bench_c2", "bench_c2_c64", "bench_c3", "bench_c5"]:
    try:
        l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(l["value"], 1), "e2e", round(l["e2e"]["value"], 1), {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in l.get("kernels", {}).items()},
              "iter", round(l.get("iteration", {}).get("ms", 0) * 1e3, 1), "cpu", (l.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "no line", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file $O/launches_default.csv \
    python bench.py --no-cpu-baseline --no-e2e --no-also --steps 1 > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_pattern -s 40 -c 1 -f -o $O/spmv_pattern_c4 \
    python bench.py --no-cpu-baseline --no-e2e --no-also --steps 1 > $O/ncu_pat_c4.log 2>&1; echo "ncu pattern c4 rc=$?"
