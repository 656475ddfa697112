#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
timeout 900 python -m pytest tests/test_gpu_cg2.py -m "gpu and not fullsize" -q > $O/r02_pytest_cg2.log 2>&1; echo "pytest cg2 rc=$?"; tail -4 $O/r02_pytest_cg2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --opt trace=256 > $O/r02_bench_n2.json 2> $O/r02_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-200 $O/r02_bench_n2.json; tail -3 $O/r02_bench_n2.err
python tools/trace_report.py $O/trace_c4_n2_r*.npy
timeout 600 python tools/kbench.py --workload c4slab8 --set cg2=1 --set cg2=1,blocks_per_sm=8 --set cg2=1,blocks_per_sm=2 > $O/r02_kbench_slab.json 2> $O/r02_kbench_slab.err; echo "kbench slab rc=$?"; cut -c1-600 $O/r02_kbench_slab.json; tail -3 $O/r02_kbench_slab.err
timeout 900 python tools/kbench.py --workload c4 --set cg2=1 --set cg2=1,blocks_per_sm=8 > $O/r02_kbench_c4.json 2> $O/r02_kbench_c4.err; echo "kbench c4 rc=$?"; cut -c1-600 $O/r02_kbench_c4.json; tail -3 $O/r02_kbench_c4.err
