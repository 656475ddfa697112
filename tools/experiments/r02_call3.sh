#!/bin/bash
# round 2, call 4 (1 GPU): ncu --set full of the two kernels of the two-kernel iteration on C4
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cg2_dir_spmv -s 20 -c 1 -f -o $O/r02_dir_spmv_c4 \
    python tools/kbench.py --workload c4 --reps 1 --set cg2=1 > $O/r02_ncu_dir.log 2>&1; echo "ncu dir rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cg2_update_r -s 20 -c 1 -f -o $O/r02_update_r_c4 \
    python tools/kbench.py --workload c4 --reps 1 --set cg2=1 > $O/r02_ncu_upd.log 2>&1; echo "ncu upd rc=$?"
ls -la $O/*.ncu-rep
