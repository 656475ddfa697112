#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 1500 python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"; cut -c1-300 $O/r02_bench_default.json; tail -3 $O/r02_bench_default.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err ) 2>&1 | grep real; cut -c1-300 $O/r02_bench_reference.json
timeout 600 python tools/kbench.py --workload c2 --set cg2=1,solver=1 --set solver=2 --set cg2=0,solver=1 > $O/r02_kbench_c2.json 2> $O/r02_kbench_c2.err; echo "kbench c2 rc=$?"; cut -c1-300 $O/r02_kbench_c2.json; tail -3 $O/r02_kbench_c2.err
python -c "import __graft_entry__ as g; g.smoke()"
