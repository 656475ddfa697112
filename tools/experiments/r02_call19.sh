#!/bin/bash
# march_conv A/B: parity of the combining consumers, then kernel timings on C4 and on one eighth of it
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
timeout 900 python -m pytest tests/test_gpu_cg2.py -q -x -k "marching or two_kernel or stores_outside or shard_entry" > $O/r02_pytest_conv.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02_pytest_conv.log
timeout 900 python tools/kbench.py --workload c4 --set march_conv=1 --set march_conv=0 > $O/r02_kbench_c4_conv.json 2> $O/r02_kbench_c4_conv.err; echo "kbench c4 rc=$?"; cat $O/r02_kbench_c4_conv.json | cut -c1-600
timeout 600 python tools/kbench.py --workload c4slab8 --set march=2,march_conv=1 --set march=2,march_conv=0 --set march=1 > $O/r02_kbench_slab_conv.json 2> $O/r02_kbench_slab_conv.err; echo "kbench slab rc=$?"; cat $O/r02_kbench_slab_conv.json | cut -c1-600
