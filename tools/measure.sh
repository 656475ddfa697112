#!/bin/bash
# Measuring call of the round (1 GPU): the default bench line and the reference arm, kernel A/B timings, the ncu launch
# list of the bench command and full captures of the kernels of the two-kernel iteration.  Everything lands in
# gpurun_out/ (r02_*); an .ncu-rep is turned into text (tools/ncu_summary.py) and deleted on the box -- gpurun brings back
# at most 64 MiB per call, three reports are more than that.  What is kept for the judge is copied to profiles/.
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
timeout 1500 python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err; echo "bench rc=$?"; cut -c1-200 $O/r02_bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "reference rc=$?"; cut -c1-200 $O/r02_bench_reference.json
timeout 900 python tools/kbench.py --workload c4 --set march=1 --set march=0 --set cg2=0 --set cg2=0,pattern=0 > $O/r02_kbench_c4.json 2> $O/r02_kbench_c4.err; echo "kbench c4 rc=$?"
timeout 600 python tools/kbench.py --workload c4slab8 --set march=1 --set march=2 --set cg2=0 > $O/r02_kbench_slab.json 2> $O/r02_kbench_slab.err; echo "kbench slab rc=$?"
timeout 600 python tools/kbench.py --workload c2 --set solver=1 --set solver=1,march=2 --set solver=1,cg2=0 > $O/r02_kbench_c2.json 2> $O/r02_kbench_c2.err; echo "kbench c2 rc=$?"
timeout 600 python tools/kbench.py --workload c2 --dtype c64 --set solver=1 --set solver=1,cg2=0 > $O/r02_kbench_c2c64.json 2> $O/r02_kbench_c2c64.err; echo "kbench c2c64 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/r02_launches_default_c4.csv \
    python bench.py --no-cpu-baseline --no-e2e --no-also --steps 1 > $O/r02_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
cap() {   # cap <kernel regex> <workload> <output stem> <traffic name> <algorithmic bytes>
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s 20 -c 1 -f -o $O/$3 \
        python tools/kbench.py --workload $2 --reps 1 > $O/$3.log 2>&1; echo "ncu $3 rc=$?"
    python tools/ncu_summary.py $O/$3.ncu-rep --name $4 --algorithmic $5 --rm >> $O/r02_traffic.jsonl
}
rm -f $O/r02_traffic.jsonl
cap cg2_dir_march c4 r02_ncu_dir_march_c4 dir_spmv 1350000000
cap cg2_update_r c4 r02_ncu_update_r_c4 update_r 648000000
cap cg2_dir_spmv c4slab8 r02_ncu_dir_spmv_slab dir_spmv_slab8 171000000
cap cg2_update_r c4slab8 r02_ncu_update_r_slab update_r_slab8 82080000
du -sh $O
