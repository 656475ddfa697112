#!/bin/bash
# 1 GPU: the whole GPU suite + the full-size tests + smoke with the code as committed
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
export CGB200_PROBLEM_CACHE=/tmp/cgb200_problems
timeout -k 5 400 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $O/r02_pytest_gpu.log
timeout -k 5 200 python -m pytest tests -m "gpu and fullsize" -x -q > $O/r02_pytest_fullsize.log 2>&1; echo "pytest fullsize rc=$?"; tail -2 $O/r02_pytest_fullsize.log
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
