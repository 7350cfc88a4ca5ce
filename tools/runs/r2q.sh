#!/bin/bash
# round 2, visit Q: full GPU suite after the OOK / generic-chain work; PSD frames-per-item A/B at 2^28 and 2^26 samples
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2q_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 $O/r2q_pytest_gpu.log
for F in 16 13 8; do
  echo "fpi $F (2^28)"; LRC_PSD_FPI=$F timeout 200 python tools/bench_kernels.py --only fft 2>/dev/null | tail -1 | cut -c1-150
  echo "fpi $F (2^26)"; LRC_PSD_FPI=$F timeout 200 python tools/bench_kernels.py --only fft --quick 2>/dev/null | tail -1 | cut -c1-150
done
