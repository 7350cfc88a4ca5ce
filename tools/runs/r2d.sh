#!/bin/bash
# round 2, visit D: fused config-3 receiver (lrc_fmrx) parity + timing, 16k overlap-save after the P1 clean-up
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_fm_resample.py -m gpu -x -q > $O/r2d_pytest_fm.log 2>&1; echo "pytest fm exit $?"; tail -n 15 $O/r2d_pytest_fm.log
(cd kpn && timeout 120 ./test_gpu_blocks > ../$O/r2d_kpn_blocks.log 2>&1; echo "kpn exit $?"; tail -n 5 ../$O/r2d_kpn_blocks.log)
timeout 300 python tools/bench_kernels.py --only fmchain > $O/r2d_fmchain.jsonl 2> $O/r2d_fmchain.err; echo "fmchain exit $?"; cut -c1-330 $O/r2d_fmchain.jsonl; tail -3 $O/r2d_fmchain.err
for TPS in 4 8 16 47; do echo "tps $TPS"; LRC_FMRX_TPS=$TPS timeout 200 python tools/bench_kernels.py --only fmchain 2>/dev/null | head -1 | cut -c1-200; done
timeout 300 python tools/fastfir16k_check.py > $O/r2d_ff16k.json 2> $O/r2d_ff16k.err; echo "16k exit $?"; cat $O/r2d_ff16k.json
timeout 600 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q -k "config3 or config5" > $O/r2d_pytest_full.log 2>&1; echo "pytest full exit $?"; tail -n 15 $O/r2d_pytest_full.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fmrx_kernel -c 1 --launch-skip 3 -o $O/r2d_fmrx_full -f \
    python tools/bench_kernels.py --only fmchain > $O/r2d_ncu_fmrx.log 2>&1; echo "ncu fmrx exit $?"
