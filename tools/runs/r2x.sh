#!/bin/bash
# round 2, visit X: chain tests with the fused/unfused policy; FIR regression (single-row TMA eligibility)
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_core.py tests/test_gpu_full_size.py tests/test_gpu_kpn.py -m gpu -x -q > $O/r2x_pytest.log 2>&1; echo "pytest exit $?"; tail -n 4 $O/r2x_pytest.log
