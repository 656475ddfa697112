#!/bin/bash
# 8 GPUs, final code: config 4 (strong scaling, timeline), config 5, parity of the sharded path.  Strict inner limits.
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout -k 5 110 $TR --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 --opt trace=256 > $O/r02_bench_n8.json 2> $O/r02_bench_n8.err; echo "bench c4 n8 rc=$?"; cut -c1-110 $O/r02_bench_n8.json
python tools/trace_report.py $O/trace_c4_n8_r0.npy $O/trace_c4_n8_r3.npy $O/trace_c4_n8_r7.npy | tee $O/r02_trace_c4_n8.txt
timeout -k 5 120 $TR --master-port 29532 bench.py --workload c5 --gpus 8 --steps 3 --warmup 3 --opt trace=256 > $O/r02_bench_c5_n8.json 2> $O/r02_bench_c5_n8.err; echo "bench c5 n8 rc=$?"; cut -c1-110 $O/r02_bench_c5_n8.json
python tools/trace_report.py $O/trace_c5_n8_r0.npy $O/trace_c5_n8_r7.npy | tee $O/r02_trace_c5_n8.txt
timeout -k 5 70 $TR --master-port 29533 tools/shard_check.py --parity-only > $O/r02_shard_check_8gpu.json 2> $O/r02_shard_check_8gpu.err; echo "shard_check 8 rc=$?"; tail -c 200 $O/r02_shard_check_8gpu.json
rm -f $O/trace_c4_n8_r[1245 6].npy $O/trace_c5_n8_r[1-6].npy
