#!/bin/bash
# round 2, visit F: fmrx with the discriminator threaded through the FIR
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_fm_resample.py -m gpu -x -q -k "fm_receiver" > $O/r2f_pytest_fm.log 2>&1; echo "pytest fm exit $?"; tail -n 5 $O/r2f_pytest_fm.log
timeout 300 python tools/bench_kernels.py --only fmchain > $O/r2f_fmchain.jsonl 2> $O/r2f_fmchain.err; echo "fmchain exit $?"; cut -c1-200 $O/r2f_fmchain.jsonl; tail -3 $O/r2f_fmchain.err
timeout 300 python tools/fastfir16k_check.py > $O/r2f_ff16k.json 2> $O/r2f_ff16k.err; echo "16k exit $?"; cat $O/r2f_ff16k.json
timeout 600 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q -k "config3" > $O/r2f_pytest_full.log 2>&1; echo "pytest full exit $?"; tail -n 5 $O/r2f_pytest_full.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fmrx_kernel -c 1 --launch-skip 3 -o $O/r2f_fmrx_full -f \
    python tools/bench_kernels.py --only fmchain > $O/r2f_ncu_fmrx.log 2>&1; echo "ncu fmrx exit $?"
