#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
: > $O/s8i_unpack_variants.txt
run() { echo "LRC_UNPACK_VARIANT=$1 LRC_UNPACK_CAP=$2" | tee -a $O/s8i_unpack_variants.txt; LRC_UNPACK_VARIANT=$1 LRC_UNPACK_CAP=$2 timeout 100 python tools/bench_kernels.py --only unpack 2>>$O/s8i.err | tee -a $O/s8i_unpack_variants.txt | cut -c1-140; }
for c in 128 256 1024 100000; do run 4 $c; run 14 $c; done
: > $O/s8i_fm_caps.txt
for c in 16 32 64 128 512; do echo "LRC_FM_CAP=$c" | tee -a $O/s8i_fm_caps.txt; LRC_FM_CAP=$c timeout 100 python tools/bench_kernels.py --only fm 2>>$O/s8i.err | grep "FM discr" | tee -a $O/s8i_fm_caps.txt | cut -c1-140; done
