#!/bin/bash
# full GPU suite, smoke, then compute-sanitizer memcheck / racecheck over the round-2 paths
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m "gpu and not fullsize" -q > $O/r02_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 $O/r02_pytest_gpu.log
timeout 600 python -m pytest tests -m "gpu and fullsize" -q -s > $O/r02_pytest_fullsize.log 2>&1; echo "pytest fullsize rc=$?"; tail -4 $O/r02_pytest_fullsize.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 480 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_cg2.py -q -x -m "gpu and not fullsize" -k "two_kernel or marching or pcg_fixed or assembled or shard_entry or bad_column or graph_is_dropped" > $O/r02_sanitizer_memcheck.log 2>&1 ) 2>&1 | grep real; echo "memcheck rc=$?"; grep "ERROR SUMMARY\|passed\|failed" $O/r02_sanitizer_memcheck.log | sort | uniq -c | tail -5
( time timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_cg2.py -q -x -m "gpu and not fullsize" -k "(marching and lap3d_40x7 and f64) or (two_kernel_iteration_matches and lap3d_slab and f64) or (pcg_fixed and poisson and f64)" > $O/r02_sanitizer_racecheck.log 2>&1 ) 2>&1 | grep real; echo "racecheck rc=$?"; grep "RACECHECK SUMMARY\|passed\|failed" $O/r02_sanitizer_racecheck.log | sort | uniq -c | tail -5
