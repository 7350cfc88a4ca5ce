#!/bin/bash
# round 2, visit BF: compute-sanitizer memcheck over the OOK kernels of the tree as shipped (quiet-block reads of d_max / d_half)
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_ook_fastfir.py -m gpu -x -q -k "edge_cases or quiet or shorter or guard" > $O/r2bf_memcheck.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/r2bf_memcheck.log | tail -6
