#!/bin/bash
# round 2, visit AN: slicer sample loads with and without L1 allocation (LRC_OOK_LD), both forms, 4096 and 512 streams
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
LRC_OOK_LD=1 timeout 300 python -m pytest tests/test_gpu_ook_fastfir.py -m gpu -x -q -k "ook" > $O/r2an_pytest.log 2>&1; echo "pytest (LD=1) exit $?"; tail -n 2 $O/r2an_pytest.log
for n in 4096 512; do for kc in 1 0; do for ld in 0 1 0 1; do echo "streams $n KC=$kc LD=$ld"; LRC_OOK_LD=$ld LRC_OOK_KC=$kc timeout 200 python tools/bench_kernels.py --only ook --ook-streams $n 2>/dev/null | tail -1 | cut -c1-120; done; done; done
