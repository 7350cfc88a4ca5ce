#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
: > $O/s8j_grid_sweep.txt
for k in 1 2 4 16 1000; do
  echo "grid multiplier $k (LRC_FFT_GRID LRC_PSD_GRID LRC_FIRU8_GRID LRC_FIR_GRID LRC_RS_GRID)" | tee -a $O/s8j_grid_sweep.txt
  LRC_FFT_GRID=$k LRC_PSD_GRID=$k LRC_FIRU8_GRID=$k LRC_FIR_GRID=$k LRC_RS_GRID=$k timeout 150 python tools/bench_kernels.py --only unpack,fir,fft,fm 2>>$O/s8j.err | tee -a $O/s8j_grid_sweep.txt | cut -c1-130
done
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 1
