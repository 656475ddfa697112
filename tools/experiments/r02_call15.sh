#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
timeout 600 python -m pytest tests/test_gpu_cg2.py -m "gpu and not fullsize" -q -k "marching" > $O/r02_pytest_march.log 2>&1; echo "pytest march rc=$?"; tail -5 $O/r02_pytest_march.log
timeout 600 python tools/kbench.py --workload c4slab8 --set march=1 --set march=0 > $O/r02_kbench_slab.json 2> $O/r02_kbench_slab.err; echo "kbench slab rc=$?"; cut -c1-460 $O/r02_kbench_slab.json; tail -3 $O/r02_kbench_slab.err
timeout 900 python tools/kbench.py --workload c4 --set march=1 --set march=0 > $O/r02_kbench_c4.json 2> $O/r02_kbench_c4.err; echo "kbench c4 rc=$?"; cut -c1-460 $O/r02_kbench_c4.json; tail -3 $O/r02_kbench_c4.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cg2_dir_march -s 20 -c 1 -f -o $O/r02_dir_march_c4 \
    python tools/kbench.py --workload c4 --reps 1 --set march=1 > $O/r02_ncu_march.log 2>&1; echo "ncu march rc=$?"
