#!/bin/bash
# round 2, visit AA: OOK guard path (lone [0.0] burst) parity, branchless matcher; per-kernel times of the OOK chain
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_ook_fastfir.py tests/test_gpu_full_size.py tests/test_gpu_kpn.py -m gpu -x -q -k "ook or envelope or eat or apps" > $O/r2aa_pytest.log 2>&1; echo "pytest exit $?"; tail -n 4 $O/r2aa_pytest.log
for i in 1 2; do timeout 200 python tools/bench_kernels.py --only ook 2>/dev/null | tail -1 | cut -c1-200; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2aa_launches_ook.csv python tools/bench_kernels.py --only ook > $O/r2aa_ncu_launch.log 2>&1; echo "ncu launches exit $?"
python - <<'PY'
import csv,collections,statistics
d=collections.defaultdict(list)
rows=[r for r in csv.reader(open('gpurun_out/r2aa_launches_ook.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[1:]:
    if 'ook_' in r[ki]: d[r[ki].split('(')[0]].append(float(r[vi].replace(',','')))
for k,v in d.items(): print(k, len(v), 'median us', statistics.median(v)/1000)
PY
