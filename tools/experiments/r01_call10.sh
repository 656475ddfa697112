#!/bin/bash
set -u
O=gpurun_out
export PYTHONUNBUFFERED=1
B="timeout 100 python bench.py --no-cpu-baseline --no-e2e --no-also --steps 3"
run() { local name=$1; shift
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
run c3_p0 --workload c3 --opt pdl=0
run c3_p1 --workload c3
run c5_4b --workload c5
