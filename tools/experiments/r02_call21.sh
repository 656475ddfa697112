#!/bin/bash
# the uniform-slot pair kernel: parity, then kernel timings on C4, one eighth of it, and C2
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
timeout 900 python -m pytest tests/test_gpu_cg2.py -q -x -k "marching or two_kernel or stores_outside or shard_entry" > $O/r02_pytest_uni.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02_pytest_uni.log
timeout 900 python tools/kbench.py --workload c4 --set uni=1 --set uni=0 > $O/r02_kbench_c4_uni.json 2> $O/r02_kbench_c4_uni.err; echo "kbench c4 rc=$?"; cut -c1-420 $O/r02_kbench_c4_uni.json
timeout 600 python tools/kbench.py --workload c4slab8 --set uni=1 --set uni=0,march=2 --set uni=0,march=1 > $O/r02_kbench_slab_uni.json 2> $O/r02_kbench_slab_uni.err; echo "kbench slab rc=$?"; cut -c1-420 $O/r02_kbench_slab_uni.json
timeout 600 python tools/kbench.py --workload c2 --set solver=1,uni=1 --set solver=1,uni=0 > $O/r02_kbench_c2_uni.json 2> $O/r02_kbench_c2_uni.err; echo "kbench c2 rc=$?"; cut -c1-420 $O/r02_kbench_c2_uni.json
