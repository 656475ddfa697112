#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_cg2.py -m "gpu and not fullsize" -q > $O/r02_pytest_cg2.log 2>&1; echo "pytest cg2 rc=$?"; tail -12 $O/r02_pytest_cg2.log
timeout 1200 python -m pytest tests/test_gpu_cg2.py -m "gpu and fullsize" -q -s > $O/r02_pytest_fullsize.log 2>&1; echo "pytest fullsize rc=$?"; tail -15 $O/r02_pytest_fullsize.log
