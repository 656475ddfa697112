#!/bin/bash
# round 2 (8 GPUs): parity of the sharded path on 8 and 4 ranks, bench lines with the kernels' own timeline
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/shard_check.py --parity-only > $O/r02_shard_check_8gpu.json 2> $O/r02_shard_check_8gpu.err; echo "shard_check 8 rc=$?"; tail -c 900 $O/r02_shard_check_8gpu.json; grep -i "error\|assert\|Traceback" $O/r02_shard_check_8gpu.err | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 tools/shard_check.py --parity-only > $O/r02_shard_check_4gpu.json 2> $O/r02_shard_check_4gpu.err; echo "shard_check 4 rc=$?"; tail -c 300 $O/r02_shard_check_4gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 --opt trace=256 > $O/r02_bench_n8.json 2> $O/r02_bench_n8.err; echo "bench n8 rc=$?"; cut -c1-200 $O/r02_bench_n8.json; tail -3 $O/r02_bench_n8.err
python tools/trace_report.py $O/trace_c4_n8_r*.npy
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 5 --warmup 3 --opt trace=256 > $O/r02_bench_n4.json 2> $O/r02_bench_n4.err; echo "bench n4 rc=$?"; cut -c1-200 $O/r02_bench_n4.json
python tools/trace_report.py $O/trace_c4_n4_r*.npy
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus 8 --steps 5 --warmup 3 --opt cg2=0 > $O/r02_bench_n8_cg2off.json 2> $O/r02_bench_n8_cg2off.err; echo "bench n8 cg2=0 rc=$?"; cut -c1-200 $O/r02_bench_n8_cg2off.json
