#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
timeout 300 python tools/experiments/r02_diag_powerlaw.py 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_cg2.py -m "gpu and not fullsize" -q > $O/r02_pytest_cg2.log 2>&1; echo "pytest cg2 rc=$?"; tail -12 $O/r02_pytest_cg2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --opt trace=256 > $O/r02_bench_n2.json 2> $O/r02_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-200 $O/r02_bench_n2.json; tail -3 $O/r02_bench_n2.err
python tools/trace_report.py $O/trace_c4_n2_r*.npy
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --opt trace=256 --opt pdl=3 > $O/r02_bench_n2_pdl3.json 2> $O/r02_bench_n2_pdl3.err; echo "bench n2 pdl3 rc=$?"; cut -c1-200 $O/r02_bench_n2_pdl3.json
python tools/trace_report.py $O/trace_c4_n2_r*.npy
