#!/bin/bash
# round 2, visit AP: compute-sanitizer over the OOK kernels of the last tree (memcheck on the edge / guard / short-capture tests through
# both slicer forms, racecheck on two small cases for the trigger kernel's shared-memory pipeline)
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_ook_fastfir.py -m gpu -x -q -k "edge_cases or guard or shorter or overflow" > $O/r2ap_memcheck.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/r2ap_memcheck.log | tail -8
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_ook_fastfir.py -m gpu -x -q -k "ragged_blocks or burst_at_end" > $O/r2ap_racecheck.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $O/r2ap_racecheck.log | tail -8
