#!/bin/bash
# round 2, visit B: the 128-bit-shared-access version of the 16384-point overlap-save kernel
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_ook_fastfir.py tests/test_gpu_core.py -m gpu -x -q -k "fastfir or tight or never_reads" > $O/r2b_pytest.log 2>&1
echo "pytest exit $?"; tail -n 5 $O/r2b_pytest.log
timeout 300 python tools/fastfir16k_check.py > $O/r2b_ff16k.json 2> $O/r2b_ff16k.err; echo "16k exit $?"; cat $O/r2b_ff16k.json; tail -3 $O/r2b_ff16k.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fastfir16k -c 1 --launch-skip 4 -o $O/r2b_ff16k_full -f \
    python tools/fastfir16k_check.py > $O/r2b_ncu_ff16k.log 2>&1; echo "ncu 16k exit $?"
