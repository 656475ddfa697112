#!/bin/bash
# 4 GPUs: config 4 with a timeline, and the SM clocks / power of every GPU sampled while it runs (is the rank-dependent
# dir_spmv time of the 8-GPU run a clock difference between the GPUs?)
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv,noheader -lms 250 > $O/r02_smi_n4.csv 2>/dev/null &
SMI=$!
timeout -k 5 110 $TR --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 5 --opt trace=256 > $O/r02_bench_n4.json 2> $O/r02_bench_n4.err; echo "bench c4 n4 rc=$?"; cut -c1-110 $O/r02_bench_n4.json
kill $SMI
for r in 0 1 2 3; do python tools/trace_report.py $O/trace_c4_n4_r$r.npy; done | tee $O/r02_trace_c4_n4.txt
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_smi_n4.csv")) if len(r) >= 4]
by = collections.defaultdict(list)
for r in rows:
    try: by[int(r[0])].append((float(r[1].split()[0]), float(r[2].split()[0]), float(r[3]), r[4].strip()))
    except Exception: pass
for g, v in sorted(by.items()):
    busy = [x for x in v if x[1] > 300]
    if busy: print("gpu", g, "samples under load", len(busy), "sm MHz min/median/max", min(x[0] for x in busy), sorted(x[0] for x in busy)[len(busy)//2], max(x[0] for x in busy), "power max", max(x[1] for x in busy), "temp max", max(x[2] for x in busy), "reasons", sorted(set(x[3] for x in busy)))
PY
