#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kpn.py tests/test_gpu_ook_fastfir.py -m gpu -x -q -k "kpn or fastfir or application or cpp" > $O/r2l_pytest.log 2>&1; echo "pytest exit $?"; tail -n 6 $O/r2l_pytest.log
timeout 300 python tools/fastfir16k_check.py > $O/r2l_ff16k.json 2> $O/r2l_ff16k.err; echo "16k exit $?"; cat $O/r2l_ff16k.json; tail -2 $O/r2l_ff16k.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fastfir16k -c 1 --launch-skip 4 -o $O/r2l_ff16k_full -f \
    python tools/fastfir16k_check.py > $O/r2l_ncu_ff16k.log 2>&1; echo "ncu 16k exit $?"
