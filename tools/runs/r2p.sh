#!/bin/bash
# round 2, visit P: generic fused chain instances -- parity, then fused vs unfused timing
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_core.py -m gpu -x -q -k "chain" > $O/r2p_pytest.log 2>&1; echo "pytest exit $?"; tail -n 15 $O/r2p_pytest.log
timeout 600 python tools/bench_kernels.py --only chaing > $O/r2p_chaing.jsonl 2> $O/r2p_chaing.err; echo "bench exit $?"; cut -c1-330 $O/r2p_chaing.jsonl; tail -3 $O/r2p_chaing.err
