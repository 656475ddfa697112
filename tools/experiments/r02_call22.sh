#!/bin/bash
# balanced contiguous run partition of the marching kernel: parity, then C4 / slab timings
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
timeout 900 python -m pytest tests/test_gpu_cg2.py -q -x -k "marching or two_kernel or stores_outside or shard_entry" > $O/r02_pytest_bal.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02_pytest_bal.log
timeout 900 python tools/kbench.py --workload c4 --set march=1 --set march=1,march_lz=24 > $O/r02_kbench_c4_bal.json 2> $O/r02_kbench_c4_bal.err; echo "kbench c4 rc=$?"; cut -c1-420 $O/r02_kbench_c4_bal.json
timeout 600 python tools/kbench.py --workload c4slab8 --set march=2 --set march=2,march_lz=13 --set march=1 > $O/r02_kbench_slab_bal.json 2> $O/r02_kbench_slab_bal.err; echo "kbench slab rc=$?"; cut -c1-420 $O/r02_kbench_slab_bal.json
timeout 600 python tools/kbench.py --workload c2 --set solver=1,march=2 --set solver=1,march=1 > $O/r02_kbench_c2_bal.json 2> $O/r02_kbench_c2_bal.err; echo "kbench c2 rc=$?"; cut -c1-420 $O/r02_kbench_c2_bal.json
