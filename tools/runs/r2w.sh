#!/bin/bash
# round 2, visit W: generic fused chain vs the unfused path now that the unfused FIR has tile instances too
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 900 python tools/bench_kernels.py --only chaing > $O/r2w_chaing.jsonl 2> $O/r2w_chaing.err; echo "bench exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/r2w_chaing.jsonl'):
    j=json.loads(l); print(j['kernel'], round(j['Msamples/s']), 'Ms/s frac_hbm', round(j['frac_hbm'],3), 'TF', round(j['TFLOP/s'],1), 'unfused_ms', round(j.get('unfused_ms',0),3), 'fused_ms', round(j['ms'],3), 'speedup', round(j.get('speedup_vs_unfused',0),2))
PY
tail -3 $O/r2w_chaing.err
