#!/bin/bash
# round 2, visit BD: quiet-block boundary test (block maximum == max/2) through both slicer forms + the OOK file
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_gpu_ook_fastfir.py -m gpu -x -q -k "ook" > $O/r2bd_pytest.log 2>&1; echo "pytest exit $?"; tail -n 3 $O/r2bd_pytest.log
