#!/bin/bash
# round 2, visit N: OOK block-sum kernel v2 (folded table + TMA swizzled slabs): parity for every variant, A/B timing, launch list, ncu
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
for V in 1 2 3 0; do
  LRC_OOK_KA=$V timeout 400 python -m pytest tests/test_gpu_ook_fastfir.py tests/test_gpu_full_size.py -m gpu -x -q -k "ook or envelope" > $O/r2n_pytest_ka$V.log 2>&1; echo "pytest ka=$V exit $?"; tail -n 3 $O/r2n_pytest_ka$V.log
  LRC_OOK_KA=$V timeout 200 python tools/bench_kernels.py --only ook 2>/dev/null | tail -1 | cut -c1-200
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2n_launches_ook.csv python tools/bench_kernels.py --only ook > $O/r2n_ncu_launch.log 2>&1; echo "ncu launches exit $?"
grep -E "ook_" $O/r2n_launches_ook.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -20
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ook_block_tma -c 1 --launch-skip 2 -o $O/r2n_ookA_full -f \
    python tools/bench_kernels.py --only ook > $O/r2n_ncu_ookA.log 2>&1; echo "ncu ookA exit $?"
