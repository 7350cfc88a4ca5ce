#!/bin/bash
# round 2, visit M: PSD variants A/B (default vs 17/18/19) + parity of the new variant
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
for V in 0 17 18 19; do echo "psd variant $V"; LRC_PSD_VARIANT=$V timeout 200 python tools/bench_kernels.py --only fft 2>/dev/null | tail -1 | cut -c1-170; done
LRC_PSD_VARIANT=17 timeout 300 python -m pytest tests/test_gpu_core.py tests/test_gpu_full_size.py -m gpu -x -q -k "psd" > $O/r2m_pytest_psd17.log 2>&1; echo "pytest psd17 exit $?"; tail -3 $O/r2m_pytest_psd17.log
LRC_PSD_VARIANT=17 timeout 300 ncu --set full --clock-control none --import-source on -k regex:psd1024_warp_sacc -c 1 --launch-skip 3 -o $O/r2m_psd17_full -f \
    python tools/bench_kernels.py --only fft > $O/r2m_ncu_psd.log 2>&1; echo "ncu psd exit $?"
