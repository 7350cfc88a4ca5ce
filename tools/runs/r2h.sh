#!/bin/bash
# round 2, visit H: fmrx occupancy variants (no spills), threaded discriminator
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_fm_resample.py -m gpu -x -q -k "fm_receiver" > $O/r2h_pytest_fm.log 2>&1; echo "pytest fm exit $?"; tail -n 3 $O/r2h_pytest_fm.log
for OCC in 3 2; do echo "occ $OCC"; LRC_FMRX_OCC=$OCC timeout 200 python tools/bench_kernels.py --only fmchain 2>/dev/null | head -1 | cut -c1-200; done
LRC_FMRX_OCC=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fmrx_kernel -c 1 --launch-skip 3 -o $O/r2h_fmrx3_full -f \
    python tools/bench_kernels.py --only fmchain > $O/r2h_ncu_fmrx.log 2>&1; echo "ncu fmrx exit $?"
