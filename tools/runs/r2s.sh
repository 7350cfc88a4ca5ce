#!/bin/bash
# round 2, visit S: generic-shape FIR tile kernel: parity (incl. bit-equality with the kernel it replaces), timing
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_core.py -m gpu -x -q -k "fir or chain" > $O/r2s_pytest.log 2>&1; echo "pytest exit $?"; tail -n 6 $O/r2s_pytest.log
timeout 600 python tools/bench_kernels.py --only firg > $O/r2s_firg.jsonl 2> $O/r2s_firg.err; echo "bench exit $?"; cut -c1-120 $O/r2s_firg.jsonl; python - <<'PY'
import json
for l in open('gpurun_out/r2s_firg.jsonl'):
    j=json.loads(l); print(j['kernel'], round(j['Msamples/s']), 'Ms/s frac_hbm', round(j['frac_hbm'],3), 'TF', round(j['TFLOP/s'],1), 'speedup', round(j['speedup'],1))
PY
tail -3 $O/r2s_firg.err
