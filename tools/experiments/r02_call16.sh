#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
timeout 900 python -m pytest tests/test_gpu_cg2.py -m "gpu and not fullsize" -q > $O/r02_pytest_cg2.log 2>&1; echo "pytest cg2 rc=$?"; tail -4 $O/r02_pytest_cg2.log
timeout 600 python tools/kbench.py --workload c4slab8 --set march=1 --set march=0 --set march=1,march_lz=5 --set march=1,march_lz=7 > $O/r02_kbench_slab.json 2> $O/r02_kbench_slab.err; echo "kbench slab rc=$?"; cut -c1-460 $O/r02_kbench_slab.json; tail -3 $O/r02_kbench_slab.err
timeout 900 python tools/kbench.py --workload c4 --set march=1 --set march=0 > $O/r02_kbench_c4.json 2> $O/r02_kbench_c4.err; echo "kbench c4 rc=$?"; cut -c1-460 $O/r02_kbench_c4.json; tail -3 $O/r02_kbench_c4.err
timeout 600 python tools/kbench.py --workload c2 --set march=1,solver=1 --set march=0,solver=1 > $O/r02_kbench_c2.json 2> $O/r02_kbench_c2.err; echo "kbench c2 rc=$?"; cut -c1-460 $O/r02_kbench_c2.json; tail -3 $O/r02_kbench_c2.err
timeout 600 python tools/kbench.py --workload c2 --dtype c64 --set march=1,solver=1 --set march=0,solver=1 --set cg2=0,solver=1 > $O/r02_kbench_c2c64.json 2> $O/r02_kbench_c2c64.err; echo "kbench c2c64 rc=$?"; cut -c1-300 $O/r02_kbench_c2c64.json
