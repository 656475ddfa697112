#!/bin/bash
# march_conv: is the cross-proxy fence what costs?  + one full ncu capture of the combining kernel on C4
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
timeout 900 python tools/kbench.py --workload c4 --set march_conv=1,march_conv_fence=0 --set march_conv=1,march_conv_fence=1 --set march_conv=0 > $O/r02_kbench_c4_conv2.json 2> $O/r02_kbench_c4_conv2.err; echo "kbench c4 rc=$?"; cut -c1-420 $O/r02_kbench_c4_conv2.json
timeout 600 python tools/kbench.py --workload c4slab8 --set march=2,march_conv=1,march_conv_fence=0 --set march=2,march_conv=0 > $O/r02_kbench_slab_conv2.json 2> $O/r02_kbench_slab_conv2.err; echo "kbench slab rc=$?"; cut -c1-420 $O/r02_kbench_slab_conv2.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cg2_dir_march -s 20 -c 1 -f -o $O/r02_dir_march_conv_c4 \
    python tools/kbench.py --workload c4 --reps 1 > $O/r02_ncu_march_conv.log 2>&1; echo "ncu march rc=$?"
