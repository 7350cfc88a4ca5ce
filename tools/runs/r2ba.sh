#!/bin/bash
# round 2, visit BA: quiet blocks (block maximum <= the burst's max/2) are not read by the slicer, in either form
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_gpu_ook_fastfir.py tests/test_gpu_full_size.py tests/test_gpu_kpn.py -m gpu -x -q -k "ook or envelope or eat or apps" > $O/r2ba_pytest.log 2>&1; echo "pytest exit $?"; tail -n 2 $O/r2ba_pytest.log
for n in 4096 2048 512; do for kc in 1 0; do echo "streams $n KC=$kc"; LRC_OOK_KC=$kc timeout 200 python tools/bench_kernels.py --only ook --ook-streams $n 2>/dev/null | tail -1 | cut -c1-120; done; done
for kc in 0 1; do
LRC_OOK_KC=$kc timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -c 90 --csv --log-file $O/r2ba_launches_ook_4096_kc$kc.csv python tools/bench_kernels.py --only ook > $O/r2ba_ncu_launch_kc$kc.log 2>&1; echo "ncu launches exit $?"
python - $kc <<'PY'
import csv,collections,statistics,sys
d=collections.defaultdict(list)
rows=[r for r in csv.reader(open('gpurun_out/r2ba_launches_ook_4096_kc%s.csv' % sys.argv[1])) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); mi=h.index('Metric Name')
for r in rows[1:]:
    if 'ook_' in r[ki]: d[(r[ki].split('(')[0], r[mi])].append(float(r[vi].replace(',','')))
for k,v in d.items(): print('KC', sys.argv[1], k, len(v), 'median', statistics.median(v))
PY
done
