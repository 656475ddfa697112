#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_cg2.py -m "gpu and not fullsize" -q > $O/r02_pytest_cg2.log 2>&1; echo "pytest cg2 rc=$?"; tail -4 $O/r02_pytest_cg2.log
for opt in "pdl=1" "pdl=3" "pdl=0"; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --workload c4slab4 --gpus 2 --steps 5 --warmup 3 --opt trace=256 --opt $opt > $O/r02_bench_slab4_$opt.json 2> $O/r02_bench_slab4_$opt.err; echo "bench slab4 $opt rc=$?"; cut -c1-120 $O/r02_bench_slab4_$opt.json
python tools/trace_report.py $O/trace_c4slab4_n2_r*.npy
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29525 bench.py --workload c5 --gpus 2 --steps 3 --warmup 3 > $O/r02_bench_c5_n2.json 2> $O/r02_bench_c5_n2.err; echo "bench c5 n2 rc=$?"; cut -c1-300 $O/r02_bench_c5_n2.json; tail -3 $O/r02_bench_c5_n2.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29526 bench.py --gpus 2 --steps 3 --warmup 3 --mode rhs-split --no-cpu-baseline > $O/r02_bench_rhs_split_n2.json 2> $O/r02_bench_rhs_split_n2.err; echo "bench rhs-split n2 rc=$?"; cut -c1-300 $O/r02_bench_rhs_split_n2.json; tail -3 $O/r02_bench_rhs_split_n2.err | cut -c1-300
