#!/bin/bash
# round 2, visit AT (last tree of the round): full parity suite on the last tree (OOK edge cases through both slicer forms in one process), smoke, bench line
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > $O/r2bc_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 3 $O/r2bc_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python bench.py --no-cpu > $O/r2bc_bench_n1.json 2> $O/r2bc_bench_n1.err; echo "bench exit $?"; cut -c1-200 $O/r2bc_bench_n1.json
