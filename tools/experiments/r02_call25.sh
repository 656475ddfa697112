#!/bin/bash
# 2 GPUs, config 5 with the L2 prefetch of the halo
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29522 bench.py --workload c5 --gpus 2 --steps 2 --warmup 3 --no-e2e --opt trace=256 > $O/r02_bench_c5_n2_pf.json 2> $O/r02_bench_c5_n2_pf.err; echo "bench c5 n2 rc=$?"; cut -c1-110 $O/r02_bench_c5_n2_pf.json
python tools/trace_report.py $O/trace_c5_n2_r*.npy | tee $O/r02_trace_c5_n2_prefetch.txt
