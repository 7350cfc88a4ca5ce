#!/bin/bash
# round 2, visit R: fused FM receiver with the R2 = 8 resampler phase: parity, timing, ncu
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_fm_resample.py tests/test_gpu_full_size.py tests/test_gpu_kpn.py -m gpu -x -q -k "fm or receiver or config3 or blocks" > $O/r2r_pytest.log 2>&1; echo "pytest exit $?"; tail -n 4 $O/r2r_pytest.log
for i in 1 2; do timeout 300 python tools/bench_kernels.py --only fmchain 2>/dev/null | tail -2 | cut -c1-260; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fmrx_kernel -c 1 --launch-skip 3 -o $O/r2r_fmrx_full -f \
    python tools/bench_kernels.py --only fmchain > $O/r2r_ncu_fmrx.log 2>&1; echo "ncu fmrx exit $?"
