#!/bin/bash
# round 2 (2 GPUs): sharded parity (two-kernel iteration over NVLink), PCG tests, bench at N=2 and N=1
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_cg2.py -m "gpu and not fullsize" -q -x > $O/r02_pytest_cg2.log 2>&1; echo "pytest cg2 rc=$?"; tail -12 $O/r02_pytest_cg2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_check.py 300 256 > $O/r02_shard_check_2gpu.json 2> $O/r02_shard_check_2gpu.err; echo "shard_check rc=$?"; tail -c 1500 $O/r02_shard_check_2gpu.json; grep -v "^\[rank" $O/r02_shard_check_2gpu.err | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r02_bench_n2.json 2> $O/r02_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-1500 $O/r02_bench_n2.json; tail -5 $O/r02_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --opt cg2=0 > $O/r02_bench_n2_cg2off.json 2> $O/r02_bench_n2_cg2off.err; echo "bench n2 cg2=0 rc=$?"; cut -c1-400 $O/r02_bench_n2_cg2off.json
