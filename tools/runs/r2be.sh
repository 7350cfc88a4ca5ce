#!/bin/bash
# round 2, visit BE: last sanity of the rebuilt library (smoke, a short bench line with the extra block, the quiet-block tests)
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 200 python -m pytest tests/test_gpu_ook_fastfir.py -m gpu -x -q -k "quiet or guard or known_packets" 2>&1 | tail -n 1
timeout 300 python bench.py --no-cpu --steps 10 > $O/r2be_bench_n1.json 2> $O/r2be_bench_n1.err; echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r2be_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['extra']['config4_ook_4096_streams']['value'], d['extra']['config4_ook_4096_streams']['workload'][-70:])"
