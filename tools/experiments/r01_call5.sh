#!/bin/bash
# gpurun call 5 (2 GPUs): pattern-dictionary SpMV, LL all-reduce; single-GPU benches, then the 2-GPU sharded checks.
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
B="timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-also --steps 3"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest16.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest16.log
tail -3 $O/pytest16.log
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
run c4_pat --workload c4
run c2_pat --workload c2
run c2c64_pat --workload c2 --dtype c64
run slab_pat --workload c4slab8
run c5_6b --workload c5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/shard_check.py 300 256 > $O/shard2_s2.json 2> $O/shard2_s2.err; echo "shard_check rc=$?"
tail -c 1500 $O/shard2_s2.json
T="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --steps 3"
$T --opt trace=256 > $O/n2_c4_pat.json 2> $O/n2_c4_pat.err; echo "n2 rc=$?"
python tools/trace_report.py $O/trace_c4_n2_r0.npy $O/trace_c4_n2_r1.npy
$T --opt pattern=0 > $O/n2_c4_csr.json 2> $O/n2_c4_csr.err; echo "n2 csr rc=$?"
for f in n2_c4_pat n2_c4_csr; do python - $O/$f.json $f <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(l["value"], 1), "it/s", {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in l["kernels"].items()},
          "iter_us", round(l["iteration"]["ms"] * 1e3, 1), "e2e", round(l["e2e"]["value"], 1))
except Exception as e:
    print(sys.argv[2], "no line", e)
PY
done
