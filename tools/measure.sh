#!/bin/bash
# Measuring call of the round (1 GPU): the guard-zone test, the default bench line and the reference arm, kernel A/B
# timings, the ncu launch list of the bench command and full captures of the kernels of the two-kernel iteration.
# Everything lands in gpurun_out/ (r02_*); the summaries kept for the judge are copied to profiles/ afterwards.
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
timeout 600 python -m pytest tests/test_gpu_cg2.py -q -k "stores_outside" > $O/r02_pytest_guard.log 2>&1; echo "guard test rc=$?"; tail -3 $O/r02_pytest_guard.log
timeout 1500 python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err; echo "bench rc=$?"; cut -c1-200 $O/r02_bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "reference rc=$?"; cut -c1-200 $O/r02_bench_reference.json
timeout 900 python tools/kbench.py --workload c4 --set march=1 --set march=0 --set cg2=0 --set cg2=0,pattern=0 > $O/r02_kbench_c4.json 2> $O/r02_kbench_c4.err; echo "kbench c4 rc=$?"
timeout 600 python tools/kbench.py --workload c4slab8 --set march=1 --set march=2 --set cg2=0 > $O/r02_kbench_slab.json 2> $O/r02_kbench_slab.err; echo "kbench slab rc=$?"
timeout 600 python tools/kbench.py --workload c2 --set solver=1 --set solver=1,march=2 --set solver=1,cg2=0 > $O/r02_kbench_c2.json 2> $O/r02_kbench_c2.err; echo "kbench c2 rc=$?"
timeout 600 python tools/kbench.py --workload c2 --dtype c64 --set solver=1 --set solver=1,cg2=0 > $O/r02_kbench_c2c64.json 2> $O/r02_kbench_c2c64.err; echo "kbench c2c64 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/r02_launches_default_c4.csv \
    python bench.py --no-cpu-baseline --no-e2e --no-also --steps 1 > $O/r02_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cg2_dir_march -s 20 -c 1 -f -o $O/r02_dir_march_c4 \
    python tools/kbench.py --workload c4 --reps 1 > $O/r02_ncu_march.log 2>&1; echo "ncu march rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cg2_update_r -s 20 -c 1 -f -o $O/r02_update_r_c4 \
    python tools/kbench.py --workload c4 --reps 1 > $O/r02_ncu_upd.log 2>&1; echo "ncu update_r rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cg2_dir_spmv -s 20 -c 1 -f -o $O/r02_dir_spmv_slab \
    python tools/kbench.py --workload c4slab8 --reps 1 > $O/r02_ncu_dir_slab.log 2>&1; echo "ncu dir_spmv slab rc=$?"
