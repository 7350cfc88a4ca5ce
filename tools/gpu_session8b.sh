#!/bin/bash
# Second visit: why is the packed resampler slower?  f32x2 micro-benchmark with the FIR operand forms, and
# ncu --set full of the packed u8 FIR, the packed two-tile resampler and the scalar resampler.
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== f32x2 ubench"
timeout 60 tools/ubench/bin/f32x2 | tee $O/s8_f32x2_ubench.txt

NCU="ncu --set full --clock-control none --import-source on --kernel-name-base mangled"
echo "== ncu u8 FIR packed"
timeout 200 $NCU -k regex:fir_tile_kernelILi64ELi10ELi7ELi128ELb1 -c 2 -o $O/s8_fir_u8_packed -f \
    python tools/bench_kernels.py --quick --only fir > $O/s8_ncu_fir.log 2>&1; echo "exit $?"
echo "== ncu resampler packed"
timeout 200 $NCU -k regex:resample_dec2_kernel -c 2 -o $O/s8_rs_dec2 -f \
    python tools/bench_kernels.py --quick --only fm > $O/s8_ncu_rs2.log 2>&1; echo "exit $?"
echo "== ncu resampler scalar"
LRC_RS_VARIANT=0 timeout 200 $NCU -k regex:resample_dec_kernel -c 2 -o $O/s8_rs_dec1 -f \
    python tools/bench_kernels.py --quick --only fm > $O/s8_ncu_rs1.log 2>&1; echo "exit $?"
ls -la $O/*.ncu-rep
