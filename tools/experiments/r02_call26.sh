#!/bin/bash
# 2 GPUs: the balanced per-non-zero SpMV on row-block shards (config 5): parity, bench line, timeline.
# Inner limits well below gpurun's: a deadlocked pair of GPUs must not eat the budget again.
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -k 5 120 $TR --master-port 29521 tools/shard_check.py --parity-only > $O/r02_shard_check_2gpu.json 2> $O/r02_shard_check_2gpu.err; echo "shard_check 2 rc=$?"; tail -c 200 $O/r02_shard_check_2gpu.json; grep -i "error\|assert\|Traceback" $O/r02_shard_check_2gpu.err | head -8
timeout -k 5 150 $TR --master-port 29522 bench.py --workload c5 --gpus 2 --steps 3 --warmup 3 --opt trace=256 > $O/r02_bench_c5_n2.json 2> $O/r02_bench_c5_n2.err; echo "bench c5 n2 rc=$?"; cut -c1-110 $O/r02_bench_c5_n2.json; tail -3 $O/r02_bench_c5_n2.err | cut -c1-200
python tools/trace_report.py $O/trace_c5_n2_r*.npy | tee $O/r02_trace_c5_n2.txt
