#!/bin/bash
# round 2, visit C: 16384-point overlap-save kernel, 512-thread vs 256-thread instance
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
for NT in 512 256; do
  LRC_FASTFIR16K_THREADS=$NT timeout 300 python tools/fastfir16k_check.py > $O/r2c_ff16k_$NT.json 2> $O/r2c_ff16k_$NT.err; echo "16k nt=$NT exit $?"; cat $O/r2c_ff16k_$NT.json; tail -3 $O/r2c_ff16k_$NT.err
done
LRC_FASTFIR16K_THREADS=256 timeout 600 python -m pytest tests/test_gpu_ook_fastfir.py -m gpu -x -q -k "fastfir" > $O/r2c_pytest256.log 2>&1; echo "pytest256 exit $?"; tail -n 3 $O/r2c_pytest256.log
timeout 600 python -m pytest tests/test_gpu_ook_fastfir.py -m gpu -x -q -k "fastfir" > $O/r2c_pytest512.log 2>&1; echo "pytest512 exit $?"; tail -n 3 $O/r2c_pytest512.log
for NT in 512 256; do
LRC_FASTFIR16K_THREADS=$NT timeout 300 ncu --set full --clock-control none --import-source on -k regex:fastfir16k -c 1 --launch-skip 4 -o $O/r2c_ff16k_${NT}_full -f \
    python tools/fastfir16k_check.py > $O/r2c_ncu_ff16k_$NT.log 2>&1; echo "ncu 16k $NT exit $?"
done
