#!/bin/bash
# round 2, visit AY: slice kernel with a provably uniform warp index (no divergence guards around its votes)
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_ook_fastfir.py tests/test_gpu_full_size.py tests/test_gpu_kpn.py -m gpu -x -q -k "ook or envelope or eat or apps" > $O/r2ay_pytest.log 2>&1; echo "pytest exit $?"; tail -n 2 $O/r2ay_pytest.log
for n in 4096 2048 1024 512; do echo "streams $n KC=1"; LRC_OOK_KC=1 timeout 200 python tools/bench_kernels.py --only ook --ook-streams $n 2>/dev/null | tail -1 | cut -c1-120; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file $O/r2ay_launches_ook_512.csv python tools/bench_kernels.py --only ook --ook-streams 512 > $O/r2ay_ncu_launch_512.log 2>&1; echo "ncu launches exit $?"
python - <<'PY'
import csv,collections,statistics,sys
d=collections.defaultdict(list)
rows=[r for r in csv.reader(open('gpurun_out/r2ay_launches_ook_512.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[1:]:
    if 'ook_' in r[ki]: d[r[ki].split('(')[0]].append(float(r[vi].replace(',','')))
for k,v in d.items(): print(k, len(v), 'median us', statistics.median(v)/1000)
PY
