#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
: > $O/s8l_psd_fft_sweep.txt
run() { echo "$*" | tee -a $O/s8l_psd_fft_sweep.txt; env "$@" timeout 100 python tools/bench_kernels.py --only fft 2>>$O/s8l.err | tee -a $O/s8l_psd_fft_sweep.txt | cut -c1-120; }
for v in 0 10 11 12 13 14 15; do run LRC_PSD_VARIANT=$v; done
for f in 8 32 64; do run LRC_PSD_FPI=$f; done
for g in 8 32 64; do run LRC_FFT_GRID=$g LRC_PSD_GRID=$g; done
