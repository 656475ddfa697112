#!/bin/bash
# 4 GPUs: config 4 after (a) 16 halo entries in flight per lane of the pushing warp, (b) bottom-halo runs last
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout -k 5 80 $TR --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 5 --opt trace=256 > $O/r02_bench_n4.json 2> $O/r02_bench_n4.err; echo "bench c4 n4 rc=$?"; cut -c1-110 $O/r02_bench_n4.json; tail -2 $O/r02_bench_n4.err | cut -c1-200
for r in 0 1 2 3; do python tools/trace_report.py $O/trace_c4_n4_r$r.npy; done | tee $O/r02_trace_c4_n4.txt
