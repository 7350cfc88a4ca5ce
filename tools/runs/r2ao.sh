#!/bin/bash
# round 2, visit AO: slice kernel with three batches in turn on 20 warps; slicer sample loads with / without L1 allocation (LRC_OOK_LD)
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_ook_fastfir.py tests/test_gpu_full_size.py -m gpu -x -q -k "ook" > $O/r2ao_pytest.log 2>&1; echo "pytest exit $?"; tail -n 2 $O/r2ao_pytest.log
LRC_OOK_LD=1 timeout 300 python -m pytest tests/test_gpu_ook_fastfir.py -m gpu -x -q -k "ook" > $O/r2ao_pytest_ld1.log 2>&1; echo "pytest (LD=1) exit $?"; tail -n 2 $O/r2ao_pytest_ld1.log
for n in 4096 2048 512; do for kc in 1 0; do for ld in 0 1; do echo "streams $n KC=$kc LD=$ld"; LRC_OOK_LD=$ld LRC_OOK_KC=$kc timeout 200 python tools/bench_kernels.py --only ook --ook-streams $n 2>/dev/null | tail -1 | cut -c1-120; done; done; done
LRC_OOK_KC=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r2ao_launches_ook_4096.csv python tools/bench_kernels.py --only ook > $O/r2ao_ncu_launch.log 2>&1; echo "ncu launches exit $?"
python - <<'PY'
import csv,collections,statistics,sys
d=collections.defaultdict(list)
rows=[r for r in csv.reader(open('gpurun_out/r2ao_launches_ook_4096.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[1:]:
    if 'ook_' in r[ki]: d[r[ki].split('(')[0]].append(float(r[vi].replace(',','')))
for k,v in d.items(): print(k, len(v), 'median us', statistics.median(v)/1000)
PY
