#!/bin/bash
# round 2, visit G: bench.py with the extra block (N = 1), fmrx discriminator-threaded version
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_fm_resample.py -m gpu -x -q -k "fm_receiver" > $O/r2g_pytest_fm.log 2>&1; echo "pytest fm exit $?"; tail -n 5 $O/r2g_pytest_fm.log
timeout 300 python tools/bench_kernels.py --only fmchain > $O/r2g_fmchain.jsonl 2> $O/r2g_fmchain.err; echo "fmchain exit $?"; cut -c1-200 $O/r2g_fmchain.jsonl; tail -3 $O/r2g_fmchain.err
timeout 600 python bench.py > $O/r2g_bench_n1.json 2> $O/r2g_bench_n1.err; echo "bench exit $?"; tail -5 $O/r2g_bench_n1.err; python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2g_bench_n1.json").read().strip().splitlines()[-1])
    print("value", j["value"], "frac", j["roofline"]["frac"], "e2e", j["e2e"]["value"], "e2e_u8", j["e2e_u8"]["value"], "h2d", j["e2e"]["h2d_copy_gbs_per_gpu_all_ranks_copying"])
    for k, v in j.get("extra", {}).items():
        print(k, round(v["value"]), "Ms/s", round(v["ms"], 3), "ms frac", round(v["roofline"]["frac"], 3), v["roofline"]["bound"])
    print(j["run"])
except Exception as e:
    print("parse failed", e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fmrx_kernel -c 1 --launch-skip 3 -o $O/r2g_fmrx_full -f \
    python tools/bench_kernels.py --only fmchain > $O/r2g_ncu_fmrx.log 2>&1; echo "ncu fmrx exit $?"
