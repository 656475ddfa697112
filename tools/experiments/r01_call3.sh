#!/bin/bash
# gpurun call 3 (2 GPUs): tests incl. scheduled SpMM; C3 with/without the row schedule; C4 p0/p1; 2-GPU sharded trace.
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
B="timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-also --steps 3"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest14.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest14.log
tail -3 $O/pytest14.log
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
run c3_sched --workload c3
run c3_nosched --workload c3 --opt spmm_schedule=0
run c3_k8 --workload c3 --k 8
run c3_k8_nosched --workload c3 --k 8 --opt spmm_schedule=0
run c3_c64 --workload c3 --dtype f32
run c4_p0 --workload c4 --opt pdl=0
run c4_p1 --workload c4
T="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --steps 3"
$T --opt trace=256 > $O/n2_c4_trace.json 2> $O/n2_c4_trace.err; echo "n2 trace rc=$?"
python tools/trace_report.py $O/trace_c4_n2_r0.npy $O/trace_c4_n2_r1.npy
$T --opt pdl=0 > $O/n2_c4_p0.json 2> $O/n2_c4_p0.err; echo "n2 p0 rc=$?"
$T > $O/n2_c4.json 2> $O/n2_c4.err; echo "n2 rc=$?"
for f in n2_c4_trace n2_c4_p0 n2_c4; do python - $O/$f.json $f <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(l["value"], 1), "it/s", {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in l["kernels"].items()},
          "iter_us", round(l["iteration"]["ms"] * 1e3, 1), "e2e", round(l["e2e"]["value"], 1))
except Exception as e:
    print(sys.argv[2], "no line", e)
PY
done
