#!/bin/bash
# 2 GPUs, last call of the round: parity of the sharded path with three pushing warps per block, 38-plane shards timeline
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -k 3 40 $TR --master-port 29521 tools/shard_check.py --parity-only > $O/r02_shard_check_2gpu.json 2> $O/r02_shard_check_2gpu.err; echo "shard_check 2 rc=$?"; tail -c 120 $O/r02_shard_check_2gpu.json; grep -i "error\|assert\|Traceback" $O/r02_shard_check_2gpu.err | head -5
timeout -k 3 40 $TR --master-port 29523 bench.py --workload c4slab4 --gpus 2 --steps 5 --warmup 3 --no-e2e --opt trace=256 > $O/r02_bench_slab4_final.json 2> $O/r02_bench_slab4_final.err; echo "bench slab4 rc=$?"; cut -c1-110 $O/r02_bench_slab4_final.json
python tools/trace_report.py $O/trace_c4slab4_n2_r*.npy | tee $O/r02_trace_slab4_n2_final.txt
