#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kpn.py -m gpu -x -q > $O/r2k_pytest_kpn.log 2>&1; echo "pytest kpn exit $?"; tail -n 12 $O/r2k_pytest_kpn.log
