#!/bin/bash
# round 2, visit Y: ncu --set full of the OOK block-sum kernel (default variant) and the slicer
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ook_block_tma -c 1 --launch-skip 2 -o $O/r2y_ookA_full -f python tools/bench_kernels.py --only ook > $O/r2y_ncu_ookA.log 2>&1; echo "ncu A exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ook_rle -c 1 --launch-skip 2 -o $O/r2y_ookC_full -f python tools/bench_kernels.py --only ook > $O/r2y_ncu_ookC.log 2>&1; echo "ncu C exit $?"
