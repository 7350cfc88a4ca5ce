#!/bin/bash
# round 2, visit AV: ncu --set full of the OOK kernels of the tree as shipped that have no capture yet (slicer one-warp-per-stream
# with precomputed thresholds at 4096 streams; burst, scan, scatter at 512 streams)
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ook_rle -c 1 --launch-skip 2 -o $O/r2av_ookC_full -f python tools/bench_kernels.py --only ook > $O/r2av_ncu_ookC.log 2>&1; echo "ncu C exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'ook_burst|ook_scan|ook_scatter|ook_slice|ook_match' -c 5 --launch-skip 10 -o $O/r2av_ook512_full -f python tools/bench_kernels.py --only ook --ook-streams 512 > $O/r2av_ncu_ook512.log 2>&1; echo "ncu 512 exit $?"
