#!/bin/bash
# Third visit: parity of the changed kernels, then the A/B table (packed resampler with the deep fill, double-buffered u8 FIR).
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -x -q > $O/s8e_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 3 $O/s8e_pytest_gpu.log
timeout 200 python tools/bench_kernels.py --only fir,fm,fmchain > $O/s8e_kernels.jsonl 2> $O/s8e_kernels.err; cut -c1-175 $O/s8e_kernels.jsonl
echo "== scalar resampler"
LRC_RS_VARIANT=0 timeout 200 python tools/bench_kernels.py --only fm,fmchain > $O/s8e_kernels_rs_scalar.jsonl 2>> $O/s8e_kernels.err; cut -c1-175 $O/s8e_kernels_rs_scalar.jsonl
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base mangled"
timeout 200 $NCU -k regex:fir_tile_kernelILi64ELi10ELi7ELi128ELb1 -c 1 --launch-skip 4 -o $O/s8e_fir_u8 -f \
    python tools/bench_kernels.py --only fir > $O/s8e_ncu_fir.log 2>&1; echo "exit $?"
timeout 200 $NCU -k regex:resample_dec2_kernel -c 1 --launch-skip 4 -o $O/s8e_rs_dec2 -f \
    python tools/bench_kernels.py --only fm > $O/s8e_ncu_rs2.log 2>&1; echo "exit $?"
