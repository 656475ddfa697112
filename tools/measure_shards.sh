#!/bin/bash
# gpurun call (N GPUs, N = $1): sharded parity (pattern kernel with halo, LL all-reduce) and the row-block bench with a timeline.
set -u
N=${1:-4}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29512 tools/shard_check.py 96 64 > $O/shard2_s3.json 2> $O/shard2_s3.err; echo "shard_check 2 rc=$?"
tail -c 700 $O/shard2_s3.json; echo
if [ "$N" -gt 2 ]; then
$TR --nproc-per-node $N --master-port 29513 tools/shard_check.py 96 64 > $O/shard${N}_s3.json 2> $O/shard${N}_s3.err; echo "shard_check $N rc=$?"
tail -c 700 $O/shard${N}_s3.json; echo
fi
B="$TR --nproc-per-node $N --master-port 29511 bench.py --gpus $N --no-cpu-baseline --steps 3"
$B --opt trace=256 > $O/n${N}_c4_pat.json 2> $O/n${N}_c4_pat.err; echo "bench n$N rc=$?"
python tools/trace_report.py $O/trace_c4_n${N}_r0.npy $O/trace_c4_n${N}_r1.npy $O/trace_c4_n${N}_r$((N-1)).npy
$B --opt pattern=0 > $O/n${N}_c4_csr.json 2> $O/n${N}_c4_csr.err; echo "bench n$N csr rc=$?"
for f in n${N}_c4_pat n${N}_c4_csr; do python - $O/$f.json $f <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(l["value"], 1), "it/s", {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in l["kernels"].items()},
          "iter_us", round(l["iteration"]["ms"] * 1e3, 1), "e2e", round(l["e2e"]["value"], 1))
except Exception as e:
    print(sys.argv[2], "no line", e)
PY
done
