#!/bin/bash
# round 2, visit AR: ncu --set full of the trigger kernel (walker with predicate update test)
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ook_trigger -c 1 --launch-skip 2 -o $O/r2ar_ookB_full -f python tools/bench_kernels.py --only ook > $O/r2ar_ncu_ookB.log 2>&1; echo "ncu B exit $?"
